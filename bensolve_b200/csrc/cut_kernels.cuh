// sm_100a kernels of the cut (K1 classify, K2 compact, K3 edge-intersect, K4 pair adjacency,
// K5 rewire/append).  All launches of one cut are stream-ordered; every kernel sizes its work
// from the device-side CutCtl, so the host never has to read a count back between stages.
// Tensor cores are deliberately not used: the path is matrix-vector and list/bitset work.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "cut_bodies.h"

__device__ __forceinline__ u64 b200_globaltimer() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define K_THREADS 256
#define POLY_EPS_D 1e-9   // POLY_EPS, bslv_poly.h:47
#define SCAN_THREADS 1024
#define ST_SKIP_A (ST_REDUNDANT | ST_OVF_A | ST_ERR_DEGENERATE | ST_NEED_BIG)
#define ST_SKIP_B (ST_SKIP_A | ST_OVF_B)

// ------------------------------------------------------------------ block primitives
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		u32 n = __shfl_up_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) >= o) v += n;
	}
	return v;
}

// exclusive scan of one value per thread across the block; `total` = block sum (all threads)
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *warp_sums /* >= 33 */, u32 &total)
{
	const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	u32 incl = warp_incl_scan(v);
	if (lane == 31) warp_sums[wid] = incl;
	__syncthreads();
	if (wid == 0) {
		u32 s = lane < nw ? warp_sums[lane] : 0;
		u32 si = warp_incl_scan(s);
		warp_sums[lane] = si - s;
		if (lane == 31) warp_sums[32] = si;
	}
	__syncthreads();
	u32 r = incl - v + warp_sums[wid];
	total = warp_sums[32];
	__syncthreads();
	return r;
}

// three scans in one pass (same barriers as one)
__device__ __forceinline__ void block_excl_scan3(const u32 v[3], u32 *warp_sums /* >= 99 */, u32 r[3], u32 total[3])
{
	const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	u32 incl[3];
#pragma unroll
	for (int k = 0; k < 3; k++) {
		incl[k] = warp_incl_scan(v[k]);
		if (lane == 31) warp_sums[33 * k + wid] = incl[k];
	}
	__syncthreads();
	if (wid < 3) {
		u32 s = lane < nw ? warp_sums[33 * wid + lane] : 0;
		u32 si = warp_incl_scan(s);
		warp_sums[33 * wid + lane] = si - s;
		if (lane == 31) warp_sums[33 * wid + 32] = si;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 3; k++) {
		r[k] = incl[k] - v[k] + warp_sums[33 * k + wid];
		total[k] = warp_sums[33 * k + 32];
	}
	__syncthreads();
}

// ------------------------------------------------------------------ stage 0: begin
__global__ void k_begin(DevState S, CutParams P)
{
	if (threadIdx.x || blockIdx.x) return;
	*S.cur = P;
	CutCtl *c = S.ctl;
	c->status = 0;
	c->n_strict = 0;
	c->min_strict_row = B200_NONE;
	c->min_strict_slot = B200_NONE;
	c->n_zp = c->n_zp_projected = 0;
	c->n_vis = c->n_new = c->inc_new = c->padj_new = 0;
	c->n_minus = c->n_zero = 0;
	c->n_pairs = c->adj_new = c->n_dead_facets = 0;
	c->n_live_scanned = 0;
	c->n_local = c->wl = c->mpad = c->n_surv = 0;
	S.facet_cnt[P.facet] = 0;
	S.facet_alive[P.facet] = 1;
}

// begin for the device-resident batch path: the halfspace is built on the device from the dual
// point vals[i] with the default callback's meaning (cone_polar, bslv_poly.c:30-39)
__global__ void k_begin_dev(DevState S, const double *vals, const unsigned char *ideal, u64 i, u32 facet, u32 batch_first, u32 seq, u32 zp_done)
{
	if (threadIdx.x || blockIdx.x) return;
	CutParams P;
	double hh = 0;
	for (int j = 0; j < B200_MAXD; j++) {
		const double v = j < S.d ? vals[i * S.d + j] : 0.0;
		P.h[j] = v;
		hh = __dadd_rn(hh, __dmul_rn(v, v));
	}
	P.alpha = (ideal && ideal[i]) ? 0.0 : -1.0;
	for (int id = 0; id < 2; id++) {
		const double thr = id ? 0.0 : P.alpha;
		P.hi[id] = __dadd_rn(thr, POLY_EPS_D);
		P.mid[id] = __dadd_rn(thr, 1.0e-2 * POLY_EPS_D);
		P.lo[id] = __dsub_rn(thr, POLY_EPS_D);
	}
	P.hh = hh;
	P.facet = facet;
	P.batch_first = batch_first;
	P.seq = seq;
	P.zp_done = zp_done;
	*S.cur = P;
	CutCtl *c = S.ctl;
	c->status = 0;
	c->n_strict = 0;
	c->min_strict_row = B200_NONE;
	c->min_strict_slot = B200_NONE;
	c->n_zp = c->n_zp_projected = 0;
	c->n_vis = c->n_new = c->inc_new = c->padj_new = 0;
	c->n_minus = c->n_zero = 0;
	c->n_pairs = c->adj_new = c->n_dead_facets = 0;
	c->n_live_scanned = 0;
	c->n_local = c->wl = c->mpad = c->n_surv = 0;
	S.facet_cnt[facet] = 0;
	S.facet_alive[facet] = 1;
}

// ------------------------------------------------------------------ K1: classify
// One tile = B200_TILE consecutive rows.  Thread t of a block handles rows 2t,2t+1 (+512 per
// iteration) so every warp-wide load is one contiguous 512-byte run per coordinate (double2).
// Algorithmic traffic: 8*d bytes read + 1 byte written per live row.
template <int D>
__global__ void __launch_bounds__(K_THREADS) k_classify(DevState S)
{
	__shared__ u32 red[4][K_THREADS / 32];
	const CutParams &P = *S.cur;
	const int d = D > 0 ? D : S.d;
	double h[D > 0 ? D : B200_MAXD];
#pragma unroll
	for (int j = 0; j < d; j++) h[j] = P.h[j];
	const double hi0 = P.hi[0], hi1 = P.hi[1], mid0 = P.mid[0], mid1 = P.mid[1], lo0 = P.lo[0], lo1 = P.lo[1];
	const u32 nrows = S.ctl->nrows;
	const u32 ntiles = (nrows + B200_TILE - 1) / B200_TILE;
	const size_t cap = S.cap_rows;
	u32 strict_cnt = 0, zp_cnt = 0, live_cnt = 0, min_row = B200_NONE;

	for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		u32 nonplus = 0;
#pragma unroll
		for (int it = 0; it < (int)(B200_TILE / (2 * K_THREADS)); it++) {
			const u32 r = tile * B200_TILE + it * 2 * K_THREADS + 2 * threadIdx.x;   // even
			const u32 lw = S.live[r >> 5] >> (r & 31);
			const u32 iw = S.ideal[r >> 5] >> (r & 31);
			const bool l0 = lw & 1u, l1 = lw & 2u;
			u8 c0 = CLS_DEAD, c1 = CLS_DEAD;
			if (l0 | l1) {
				double t0 = 0.0, t1 = 0.0;
#pragma unroll
				for (int j = 0; j < d; j++) {
					const double2 x = *reinterpret_cast<const double2 *>(S.coord + j * cap + r);
					if (j == 0) {
						t0 = __dmul_rn(h[0], x.x);
						t1 = __dmul_rn(h[0], x.y);
					} else {
						t0 = __dadd_rn(t0, __dmul_rn(h[j], x.x));
						t1 = __dadd_rn(t1, __dmul_rn(h[j], x.y));
					}
				}
				if (l0) {
					const bool id = iw & 1u;
					const double hi = id ? hi1 : hi0, mid = id ? mid1 : mid0, lo = id ? lo1 : lo0;
					c0 = t0 > hi ? CLS_PLUS : t0 > mid ? CLS_ZP : t0 > lo ? CLS_ZERO : CLS_MINUS;
					live_cnt++;
					nonplus += (c0 != CLS_PLUS);
					zp_cnt += (c0 == CLS_ZP);
					if (t0 < lo) { strict_cnt++; min_row = min(min_row, r); }
				}
				if (l1) {
					const bool id = iw & 2u;
					const double hi = id ? hi1 : hi0, mid = id ? mid1 : mid0, lo = id ? lo1 : lo0;
					c1 = t1 > hi ? CLS_PLUS : t1 > mid ? CLS_ZP : t1 > lo ? CLS_ZERO : CLS_MINUS;
					live_cnt++;
					nonplus += (c1 != CLS_PLUS);
					zp_cnt += (c1 == CLS_ZP);
					if (t1 < lo) { strict_cnt++; min_row = min(min_row, r + 1); }
				}
			}
			*reinterpret_cast<uchar2 *>(S.cls + r) = make_uchar2(c0, c1);
		}
		// per-tile count of non-PLUS rows (input of the ordered compaction)
		nonplus = __reduce_add_sync(0xffffffffu, nonplus);
		if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = nonplus;
		__syncthreads();
		if (threadIdx.x == 0) {
			u32 s = 0;
#pragma unroll
			for (int w = 0; w < K_THREADS / 32; w++) s += red[0][w];
			S.tile_cnt[tile] = s;
		}
		__syncthreads();
	}
	// trigger bookkeeping: one atomic per block and counter, only when non-zero
	strict_cnt = __reduce_add_sync(0xffffffffu, strict_cnt);
	zp_cnt = __reduce_add_sync(0xffffffffu, zp_cnt);
	live_cnt = __reduce_add_sync(0xffffffffu, live_cnt);
	min_row = __reduce_min_sync(0xffffffffu, min_row);
	if ((threadIdx.x & 31) == 0) {
		red[0][threadIdx.x >> 5] = strict_cnt;
		red[1][threadIdx.x >> 5] = zp_cnt;
		red[2][threadIdx.x >> 5] = live_cnt;
		red[3][threadIdx.x >> 5] = min_row;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 s = 0, z = 0, l = 0, m = B200_NONE;
#pragma unroll
		for (int w = 0; w < K_THREADS / 32; w++) {
			s += red[0][w];
			z += red[1][w];
			l += red[2][w];
			m = min(m, red[3][w]);
		}
		if (s) { atomicAdd(&S.ctl->n_strict, s); atomicMin(&S.ctl->min_strict_row, m); }
		if (z) atomicAdd(&S.ctl->n_zp, z);
		if (l) atomicAdd(&S.ctl->n_live_scanned, l);
	}
}

// ------------------------------------------------------------------ K2a: scan tile counts, decide
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(DevState S)
{
	__shared__ u32 ws[33];
	CutCtl *c = S.ctl;
	const u32 ntiles = (c->nrows + B200_TILE - 1) / B200_TILE;
	u32 carry = 0;
	for (u32 base = 0; base < ntiles; base += SCAN_THREADS) {
		u32 i = base + threadIdx.x, v = i < ntiles ? S.tile_cnt[i] : 0, tot;
		u32 e = block_excl_scan(v, ws, tot);
		if (i < ntiles) S.tile_base[i] = carry + e;
		carry += tot;
	}
	if (threadIdx.x == 0) {
		c->n_vis = carry;
		if (c->n_strict == 0) {                      // nothing to cut: redundant (bslv_poly.c:132-136)
			c->status |= ST_REDUNDANT;
			S.facet_alive[S.cur->facet] = 0;
		} else
			c->min_strict_slot = S.row_slot[c->min_strict_row];
	}
}

// ------------------------------------------------------------------ K2b: ordered compaction
// Thread t owns 8 consecutive rows of the tile (one 8-byte load of class bytes); tiles without a
// non-PLUS row are skipped after reading one counter.
__global__ void __launch_bounds__(K_THREADS) k_scatter(DevState S)
{
	__shared__ u32 ws[33];
	if (S.ctl->status & (ST_SKIP_A & ~(u32)ST_REDUNDANT)) return;   // also runs for a redundant halfspace: its marks are undone below
	const u32 nrows = S.ctl->nrows;
	const u32 ntiles = (nrows + B200_TILE - 1) / B200_TILE;
	for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
		if (S.tile_cnt[tile] == 0) continue;         // block-uniform
		const u32 r0 = tile * B200_TILE + threadIdx.x * 8;
		const u64 cw = *reinterpret_cast<const u64 *>(S.cls + r0);
		u32 n = 0;
#pragma unroll
		for (int b = 0; b < 8; b++) {
			u32 c = (u32)(cw >> (8 * b)) & 0xffu;
			n += (c >= CLS_ZP && c <= CLS_MINUS) && (r0 + b < nrows);
		}
		u32 tot, e = block_excl_scan(n, ws, tot);
		u32 w = S.tile_base[tile] + e;
#pragma unroll
		for (int b = 0; b < 8; b++) {
			u32 c = (u32)(cw >> (8 * b)) & 0xffu;
			if ((c >= CLS_ZP && c <= CLS_MINUS) && (r0 + b < nrows)) S.vis[w++] = r0 + b;
		}
	}
}

// ------------------------------------------------------------------ ZERO+ closure (rare)
__global__ void __launch_bounds__(SCAN_THREADS) k_zp_closure(DevState S)
{
	__shared__ int changed;
	CutCtl *c = S.ctl;
	if ((c->status & ST_SKIP_A) || c->n_zp == 0) return;
	const CutParams P = *S.cur;
	const u32 n = c->n_vis;
	do {
		__syncthreads();
		if (threadIdx.x == 0) changed = 0;
		__syncthreads();
		for (u32 i = threadIdx.x; i < n; i += blockDim.x)
			if (zp_activate(S, P, i)) changed = 1;
		__threadfence_block();
		__syncthreads();
	} while (changed);
}

// ------------------------------------------------------------------ generic map stages
#define B200_GRID_STRIDE(i, n) \
	for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < (u64)(n); i += (u64)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(K_THREADS) k_count(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	B200_GRID_STRIDE(i, S.ctl->n_vis) count_outputs(S, (u32)i);
}

// exclusive scan of the (rows, incidence entries, PLUS neighbours) triples + capacity plan
__global__ void __launch_bounds__(SCAN_THREADS) k_scan3_plan(DevState S)
{
	__shared__ u32 ws[33];
	CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_A) return;
	const u32 n = c->n_vis;
	u32 carry[3] = {0, 0, 0};
	for (u32 base = 0; base < n; base += SCAN_THREADS) {
		u32 i = base + threadIdx.x;
#pragma unroll
		for (int k = 0; k < 3; k++) {
			u32 v = i < n ? S.cnt3[3 * (size_t)i + k] : 0, tot;
			u32 e = block_excl_scan(v, ws, tot);
			if (i < n) S.base3[3 * (size_t)i + k] = carry[k] + e;
			carry[k] += tot;
		}
	}
	if (threadIdx.x == 0) {
		c->n_new = carry[0];
		c->inc_new = carry[1];
		c->padj_new = carry[2];
		S.facet_cnt[S.cur->facet] = carry[0];          // every new row lies on the new facet
		u32 st = 0;
		if ((u64)c->nrows + carry[0] > S.cap_rows) st |= ST_OVF_ROWS;
		if ((u64)c->inc_used + carry[1] > S.cap_inc) st |= ST_OVF_INC;
		if (carry[2] > S.cap_padj) st |= ST_OVF_PADJ;
		c->status |= st;
	}
}

__global__ void __launch_bounds__(K_THREADS) k_emit(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	const CutParams &P = *S.cur;
	B200_GRID_STRIDE(i, S.ctl->n_vis) emit_outputs(S, P, (u32)i);
}

__global__ void __launch_bounds__(K_THREADS) k_dead_facets(DevState S)
{
	if (S.ctl->status & ST_REDUNDANT) {
		B200_GRID_STRIDE(i, S.ctl->n_vis) reset_class(S, (u32)i);
		return;
	}
	if (S.ctl->status & ST_SKIP_A) return;
	B200_GRID_STRIDE(i, S.ctl->n_vis) collect_dead_facets(S, (u32)i);
}

// ------------------------------------------------------------------ K4: pair adjacency on packed bitsets
__global__ void __launch_bounds__(K_THREADS) k_pairs_reset(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		S.ctl->n_pairs = 0;
		S.ctl->n_surv = 0;
		S.ctl->status &= ~(u32)ST_OVF_B;
	}
	B200_GRID_STRIDE(j, S.ctl->n_new) { S.deg[j] = 0; S.adj_fill[j] = 0; }
}

__global__ void __launch_bounds__(K_THREADS) k4_assign(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	B200_GRID_STRIDE(j, S.ctl->n_new) k4_assign_columns(S, (u32)j);
}
__global__ void k4_plan_kernel(DevState S)
{
	if (threadIdx.x || blockIdx.x) return;
	if (S.ctl->status & ST_SKIP_A) return;
	k4_plan(S);
}
__global__ void __launch_bounds__(K_THREADS) k4_zero(DevState S)
{
	if (S.ctl->status & ST_SKIP_B) return;
	B200_GRID_STRIDE(x, (u64)S.ctl->n_local * (S.ctl->mpad / 64)) k4_zero_cols(S, x);
}
__global__ void __launch_bounds__(K_THREADS) k4_build(DevState S)
{
	if (S.ctl->status & ST_SKIP_B) return;
	B200_GRID_STRIDE(j, S.ctl->n_new) k4_build_row(S, (u32)j);
}

// AND + POPC prefilter over 64x64 tiles of the pair space.  Both row tiles of the word-major bit
// matrix are staged in shared memory K4_WCH words at a time; thread (i0, j) owns the 16 pairs
// (i0 + 4k, j), so a warp reads 32 consecutive B words (conflict-free) and one broadcast A word.
#define K4_T 64
#define K4_WCH 16
#define K4_SV 512
// `early` != 0: the grid was launched with one extra block that does no pair work.  The producer grid (k_tail) has
// written the delta record's payload to mapped host memory and has completed, so that block publishes a copy of
// the control block plus the sequence number in the EARLY header: the host applies the record to its mirror while
// the pair test and the adjacency build are still running.
// tile pairs tp0, tp0 + stride, ... of one cut's pair space (all threads of the block)
// number of tile pairs (ta <= tb) of a cut with M new rows
__device__ __forceinline__ u32 k4_tile_pairs(u32 M)
{
	const u32 nt = (M + K4_T - 1) / K4_T;
	return nt * (nt + 1) / 2;
}
__device__ __forceinline__ void k4_filter_body(const DevState &S, u32 thr, u32 M, u32 wl, u32 mpad, u32 tp0, u32 stride)
{
	__shared__ u64 sa[K4_WCH][K4_T], sb[K4_WCH][K4_T];
	__shared__ u32 sva[K4_SV], svb[K4_SV], nsv, gbase;
	const u32 nt = (M + K4_T - 1) / K4_T;
	const u32 j = threadIdx.x & (K4_T - 1), i0 = threadIdx.x >> 6;
	for (u32 tp = tp0; tp < nt * (nt + 1) / 2; tp += stride) {
		u32 ta = 0, tb = tp;                         // tp-th pair of the upper triangle, row by row (block-uniform)
		while (tb >= nt - ta) { tb -= nt - ta; ta++; }
		tb += ta;
		u32 cnt[K4_T / 4];
#pragma unroll
		for (int k = 0; k < K4_T / 4; k++) cnt[k] = 0;
		for (u32 w0 = 0; w0 < wl; w0 += K4_WCH) {
			const u32 wn = min((u32)K4_WCH, wl - w0);
			for (u32 e = threadIdx.x; e < wn * K4_T; e += K_THREADS) {
				const u32 w = e / K4_T, x = e % K4_T;
				const u32 xa = ta * K4_T + x, xb = tb * K4_T + x;
				sa[w][x] = xa < mpad ? S.bits[(size_t)(w0 + w) * mpad + xa] : 0;
				sb[w][x] = xb < mpad ? S.bits[(size_t)(w0 + w) * mpad + xb] : 0;
			}
			__syncthreads();
			for (u32 w = 0; w < wn; w++) {
				const u64 bj = sb[w][j];
#pragma unroll
				for (int k = 0; k < K4_T / 4; k++) cnt[k] += __popcll(sa[w][i0 + 4 * k] & bj);
			}
			__syncthreads();
		}
		// survivors of this tile pair are collected in shared memory and published with ONE global
		// atomic per block (thousands of single-address atomics would serialise in L2)
		const u32 b = tb * K4_T + j;
		if (threadIdx.x == 0) nsv = 0;
		__syncthreads();
#pragma unroll
		for (int k = 0; k < K4_T / 4; k++) {
			const u32 a = ta * K4_T + i0 + 4 * k;
			if (a < b && b < M && cnt[k] >= thr) {
				const u32 q = atomicAdd(&nsv, 1u);
				if (q < K4_SV) { sva[q] = a; svb[q] = b; }
				else k4_push_survivor(S, a, b);           // rare overflow of the block buffer
			}
		}
		__syncthreads();
		const u32 n = min(nsv, (u32)K4_SV);
		if (threadIdx.x == 0 && n) gbase = atomicAdd(&S.ctl->n_surv, n);
		__syncthreads();
		for (u32 q = threadIdx.x; q < n; q += K_THREADS)
			if (gbase + q < S.cap_pairs) { S.surv_a[gbase + q] = sva[q]; S.surv_b[gbase + q] = svb[q]; }
		__syncthreads();
	}
}
__global__ void __launch_bounds__(K_THREADS) k4_filter(DevState S, u32 thr, int early)
{
	cudaGridDependencySynchronize();      // programmatic dependent launch: wait for the producer grid here
	if (blockIdx.x == 0 && threadIdx.x == 0) S.dbg[13] = b200_globaltimer();
	const CutCtl *c = S.ctl;
	if (early && blockIdx.x == gridDim.x - 1) {
		if (threadIdx.x < 32 && !(c->status & ST_SKIP_B)) {
			if (threadIdx.x < sizeof(CutCtl) / 4) ((volatile u32 *)(S.stage + B200_STAGE_EARLY))[threadIdx.x] = ((const u32 *)c)[threadIdx.x];
			__threadfence_system();
			__syncwarp();
			if (threadIdx.x == 0) *(volatile u32 *)(S.stage + B200_STAGE_EARLY + B200_STAGE_SEQ) = S.cur->seq;
		}
		return;
	}
	if (c->status & ST_SKIP_B) return;
	k4_filter_body(S, thr, c->n_new, c->wl, c->mpad, blockIdx.x, gridDim.x - (early ? 1u : 0u));
}

// Containment test, one warp per surviving pair: the pair is adjacent iff no third new row contains
// inc(a) & inc(b) (edge_test, bslv_poly.c:487-505).  Lanes scan 32 candidate rows per step; only the
// non-zero words of the mask are compared (a mask holds >= d-2 bits, rarely more than a few words).
#define K4_NZ 8
__device__ __forceinline__ bool k4_contain_verdict(const DevState &S, const u64 *bits, u32 a, u32 b, u32 lane, u32 M, u32 wl, u32 mpad)
{
	bool adjacent = true;
	if (S.d != 1) {
		// the non-zero words of the mask inc(a) & inc(b): the first K4_NZ in registers of every lane
		// (the cut path needs no more), up to 32 spread one per lane (dual polytope: a ridge holds many
		// vertices), beyond that every word is compared
		u64 mw[K4_NZ], lane_m = 0;
		u32 mi[K4_NZ], lane_i = 0, nz = 0;
		for (u32 w0 = 0; w0 < wl && nz <= 32; w0 += 32) {
			const u32 w = w0 + lane;
			const u64 m = w < wl ? (bits[(size_t)w * mpad + a] & bits[(size_t)w * mpad + b]) : 0;
			u32 bal = __ballot_sync(0xffffffffu, m != 0);
			while (bal) {
				const int src = __ffs(bal) - 1;
				bal &= bal - 1;
				const u64 mv = __shfl_sync(0xffffffffu, m, src);
				if (nz < K4_NZ) { mw[nz] = mv; mi[nz] = w0 + src; }
				if (nz < 32 && lane == nz) { lane_m = mv; lane_i = w0 + src; }
				nz++;
			}
		}
		// 128 candidate rows per vote (four per lane) so that four independent load chains are in flight
		for (u32 x0 = 0; x0 < M; x0 += 128) {
			bool any = false;
#pragma unroll
			for (int h = 0; h < 4; h++) {
				const u32 x = x0 + 32 * h + lane;
				bool cont = x < M && x != a && x != b;
				if (nz <= K4_NZ) {
					if (cont) {
#pragma unroll
						for (int q = 0; q < K4_NZ; q++)
							if (q < (int)nz && (bits[(size_t)mi[q] * mpad + x] & mw[q]) != mw[q]) cont = false;
					}
				} else if (nz <= 32) {
					for (u32 q = 0; q < nz; q++) {          // warp-uniform trip count: shuffles are safe
						const u64 mq = __shfl_sync(0xffffffffu, lane_m, q);
						const u32 iq = __shfl_sync(0xffffffffu, lane_i, q);
						if (cont && (bits[(size_t)iq * mpad + x] & mq) != mq) cont = false;
					}
				} else if (cont) {
					for (u32 w = 0; w < wl && cont; w++) {
						const u64 m = bits[(size_t)w * mpad + a] & bits[(size_t)w * mpad + b];
						cont = (bits[(size_t)w * mpad + x] & m) == m;
					}
				}
				any |= cont;
			}
			if (__any_sync(0xffffffffu, any)) { adjacent = false; break; }
		}
	}
	return adjacent;
}
// Containment by columns, one warp per surviving pair: lane l owns word l (+32, +64, ...) of the
// candidate set, which is the AND of the columns (rows-on-facet bitmaps) of every facet in the mask
// inc(a) & inc(b); the pair is adjacent iff nothing but a and b survives (edge_test, bslv_poly.c:487-505).
// Work per pair: |mask| column words per lane instead of a scan over all M candidate rows.
__device__ __forceinline__ bool k4_columns_verdict(const DevState &S, u32 a, u32 b, u32 lane, u32 M, u32 wl, u32 mpad)
{
	if (S.d == 1) return true;
	const u64 *tb = k4_tbits(S, wl, mpad);
	const u32 mw = mpad / 64;
	bool other = false;
	for (u32 x0 = 0; x0 < mw && !other; x0 += 32) {            // warp-uniform trip count
		const u32 xw = x0 + lane;
		u64 acc = 0;
		if (xw < mw) acc = (xw + 1) * 64 <= M ? ~(u64)0 : (M > xw * 64 ? (((u64)1 << (M - xw * 64)) - 1) : 0);
		// narrow matrices (the usual case, <= 128 facets met by the new rows): the columns of the mask are fetched
		// together -- a load per AND would make the test a chain of dependent L2 round trips; a simple vertex pair
		// has d-2 of them
		if (wl <= 2) {
			u64 m0 = S.bits[a] & S.bits[b], m1 = wl > 1 ? (S.bits[(size_t)mpad + a] & S.bits[(size_t)mpad + b]) : 0;   // same in every lane
			u64 v[8];
#pragma unroll
			for (u32 t = 0; t < 8; t++) {
				u32 col = B200_NONE;
				if (m0) { col = (u32)__ffsll((long long)m0) - 1; m0 &= m0 - 1; }
				else if (m1) { col = 64 + (u32)__ffsll((long long)m1) - 1; m1 &= m1 - 1; }
				v[t] = (col != B200_NONE && xw < mw) ? tb[(size_t)col * mw + xw] : ~(u64)0;
			}
#pragma unroll
			for (u32 t = 0; t < 8; t++) acc &= v[t];
			while (m0 | m1) {                      // more than 8 common facets (degenerate pairs)
				u32 col;
				if (m0) { col = (u32)__ffsll((long long)m0) - 1; m0 &= m0 - 1; }
				else { col = 64 + (u32)__ffsll((long long)m1) - 1; m1 &= m1 - 1; }
				if (xw < mw) acc &= tb[(size_t)col * mw + xw];
			}
		} else
			for (u32 w = 0; w < wl; w++) {
				u64 m = S.bits[(size_t)w * mpad + a] & S.bits[(size_t)w * mpad + b];   // same value in every lane
				while (m) {
					const u32 col = w * 64 + (u32)__ffsll((long long)m) - 1;
					m &= m - 1;
					if (xw < mw) acc &= tb[(size_t)col * mw + xw];
				}
			}
		if ((a >> 6) == xw) acc &= ~((u64)1 << (a & 63));
		if ((b >> 6) == xw) acc &= ~((u64)1 << (b & 63));
		other = __any_sync(0xffffffffu, acc != 0);
	}
	return !other;
}
__device__ __forceinline__ void k4_contain_warp(const DevState &S, const u64 *, u32 s, u32 lane, u32 M, u32 wl, u32 mpad)
{
	const u32 a = S.surv_a[s], b = S.surv_b[s];
	if (k4_columns_verdict(S, a, b, lane, M, wl, mpad) && lane == 0) k4_push_pair(S, a, b);
}

template <bool COLUMNS>
__device__ __forceinline__ void contain_block_rounds(const DevState &S, u32 *pra, u32 *prb, u32 &npr, u32 &pbase, u32 bid, u32 nblocks)
{
	const CutCtl *c = S.ctl;
	const u32 M = c->n_new, wl = c->wl, mpad = c->mpad, ns = c->n_surv;
	const u32 lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	// adjacent pairs of a block round are collected in shared memory; one global atomic per round
	for (u32 s0 = bid * wpb; s0 < ns; s0 += nblocks * wpb) {   // block-uniform trip count
		const u32 s = s0 + (threadIdx.x >> 5);
		bool adj = false;
		u32 a = 0, b = 0;
		if (s < ns) {
			a = S.surv_a[s];
			b = S.surv_b[s];
			adj = COLUMNS ? k4_columns_verdict(S, a, b, lane, M, wl, mpad) : k4_contain_verdict(S, S.bits, a, b, lane, M, wl, mpad);
		}
		if (threadIdx.x == 0) npr = 0;
		__syncthreads();
		if (adj && lane == 0) {
			const u32 q = atomicAdd(&npr, 1u);
			pra[q] = a;
			prb[q] = b;
			atomicAdd(&S.deg[a], 1u);
			atomicAdd(&S.deg[b], 1u);
		}
		__syncthreads();
		if (threadIdx.x == 0 && npr) pbase = atomicAdd(&S.ctl->n_pairs, npr);
		__syncthreads();
		if (threadIdx.x < npr && pbase + threadIdx.x < S.cap_pairs) {
			S.pair_a[pbase + threadIdx.x] = pra[threadIdx.x];
			S.pair_b[pbase + threadIdx.x] = prb[threadIdx.x];
		}
		__syncthreads();
	}
}
__global__ void __launch_bounds__(K_THREADS) k4_contain(DevState S)
{
	__shared__ u32 pra[K_THREADS / 32], prb[K_THREADS / 32], npr, pbase;
	cudaGridDependencySynchronize();
	if (blockIdx.x == 0 && threadIdx.x == 0) S.dbg[14] = b200_globaltimer();
	const CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_B) return;
	if (c->n_surv > S.cap_pairs) return;             // overflow is flagged by k_adj_scan
	contain_block_rounds<true>(S, pra, prb, npr, pbase, blockIdx.x, gridDim.x);
}
// K6 keeps the row-scan form: its columns are the 10^5..10^6 live vertices
__global__ void __launch_bounds__(K_THREADS) k6_contain(DevState S)
{
	__shared__ u32 pra[K_THREADS / 32], prb[K_THREADS / 32], npr, pbase;
	const CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_B) return;
	if (c->n_surv > S.cap_pairs) return;
	contain_block_rounds<false>(S, pra, prb, npr, pbase, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_adj_scan(DevState S)
{
	__shared__ u32 ws[33];
	CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_A) return;
	if (c->status & ST_OVF_BITS) return;
	if (c->n_pairs > S.cap_pairs || c->n_surv > S.cap_pairs) {
		if (threadIdx.x == 0) c->status |= ST_OVF_PAIRS;
		return;
	}
	const u32 n = c->n_new;
	u32 carry = 0;
	for (u32 base = 0; base < n; base += SCAN_THREADS) {
		u32 j = base + threadIdx.x, v = j < n ? S.new_padj_len[j] + S.deg[j] : 0, tot;
		u32 e = block_excl_scan(v, ws, tot);
		if (j < n) S.adj_base[j] = carry + e;
		carry += tot;
	}
	if (threadIdx.x == 0) {
		c->adj_new = carry;
		if ((u64)c->adj_used + carry > S.cap_adj) c->status |= ST_OVF_ADJ;
	}
}

__global__ void __launch_bounds__(K_THREADS) k_adj_place(DevState S)
{
	if (S.ctl->status & ST_SKIP_B) return;
	B200_GRID_STRIDE(j, S.ctl->n_new) adj_place(S, (u32)j);
}
__global__ void __launch_bounds__(K_THREADS) k_adj_pair_fill(DevState S)
{
	if (S.ctl->status & ST_SKIP_B) return;
	B200_GRID_STRIDE(p, S.ctl->n_pairs) adj_pair_fill(S, (u32)p);
}
__global__ void __launch_bounds__(K_THREADS) k_adj_sort(DevState S)
{
	if (S.ctl->status & ST_SKIP_B) return;
	B200_GRID_STRIDE(j, S.ctl->n_new) adj_sort(S, (u32)j);
}

// ------------------------------------------------------------------ stage last: commit the appends
__global__ void k_finish(DevState S)
{
	if (threadIdx.x || blockIdx.x) return;
	CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_B) return;
	c->n_live = c->n_live + c->n_new - (c->n_minus + c->n_zero);
	c->nrows += c->n_new;       // NOTE: from here on ctl->nrows includes the new rows
	c->slot_cnt += c->n_new;
	c->inc_used += c->inc_new;
	c->adj_used += c->adj_new;
}

// ------------------------------------------------------------------ pack the delta for the host
__global__ void __launch_bounds__(K_THREADS) k_pack_delta(DevState S, int header_only)
{
	CutCtl *c = S.ctl;
	const StageLayout L = stage_layout(*c, S.d);
	const bool fits = header_only || L.total <= S.cap_stage;
	if (blockIdx.x == 0 && threadIdx.x < sizeof(CutCtl) / 4) {
		u32 v = ((const u32 *)c)[threadIdx.x];
		if (threadIdx.x == offsetof(CutCtl, status) / 4 && !fits) v |= ST_OVF_STAGE;
		if (threadIdx.x == offsetof(CutCtl, stage_bytes) / 4) v = (u32)L.total;
		((volatile u32 *)S.stage)[threadIdx.x] = v;       // the sequence number follows in k_publish (next launch)
	}
	if (header_only || !fits || (c->status & (ST_SKIP_B))) return;
	const u64 n = (u64)c->n_new * S.d + c->n_new + c->n_vis + c->n_dead_facets;
	B200_GRID_STRIDE(e, n) pack_delta_item(S, L, e, c->nrows - c->n_new);   // k_finish has already advanced nrows
}

// ------------------------------------------------------------------ compaction of dead rows (GC)
// Host slots are never reused (bslv_poly.c never compacts either, poly_defrag :296-312 is dead
// code), but device rows are: once dead rows outnumber live ones the live rows are packed to the
// front (stable, so rows stay sorted by slot), the two pools are repacked and adjacency entries
// are renamed.  K1 then streams live coordinates only.
struct LiveBitOf {
	const u32 *live;
	__device__ u32 operator()(u32 i) const { return (live[i >> 5] >> (i & 31)) & 1u; }
};
struct LenOfOld {
	const u32 *len, *old_of;
	__device__ u32 operator()(u32 i) const { return len[old_of[i]]; }
};

// exclusive scan of f(0..n) in three launches: tile sums, scan of tile sums, tile-local scan
template <class F> __global__ void __launch_bounds__(K_THREADS) k_gscan_reduce(F f, u32 n, u32 *tile_sum)
{
	__shared__ u32 red[K_THREADS / 32];
	const u32 base = blockIdx.x * B200_TILE + threadIdx.x * 8;
	u32 s = 0;
#pragma unroll
	for (int b = 0; b < 8; b++) s += (base + b < n) ? f(base + b) : 0;
	s = __reduce_add_sync(0xffffffffu, s);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 t = 0;
		for (int w = 0; w < K_THREADS / 32; w++) t += red[w];
		tile_sum[blockIdx.x] = t;
	}
}
__global__ void __launch_bounds__(SCAN_THREADS) k_gscan_tiles(const u32 *tile_sum, u32 ntiles, u32 *tile_base, u32 *total)
{
	__shared__ u32 ws[33];
	u32 carry = 0;
	for (u32 base = 0; base < ntiles; base += SCAN_THREADS) {
		u32 i = base + threadIdx.x, v = i < ntiles ? tile_sum[i] : 0, tot;
		u32 e = block_excl_scan(v, ws, tot);
		if (i < ntiles) tile_base[i] = carry + e;
		carry += tot;
	}
	if (threadIdx.x == 0) *total = carry;
}
template <class F> __global__ void __launch_bounds__(K_THREADS) k_gscan_apply(F f, u32 n, const u32 *tile_base, u32 *out)
{
	__shared__ u32 ws[33];
	const u32 base = blockIdx.x * B200_TILE + threadIdx.x * 8;
	u32 v[8], s = 0;
#pragma unroll
	for (int b = 0; b < 8; b++) { v[b] = (base + b < n) ? f(base + b) : 0; s += v[b]; }
	u32 tot, e = block_excl_scan(s, ws, tot);
	u32 run = tile_base[blockIdx.x] + e;
#pragma unroll
	for (int b = 0; b < 8; b++) {
		if (base + b < n) out[base + b] = run;
		run += v[b];
	}
}

__global__ void __launch_bounds__(K_THREADS) k_gc_invert(const u32 *live, const u32 *remap, u32 nrows, u32 *old_of)
{
	B200_GRID_STRIDE(r, nrows)
		if ((live[r >> 5] >> (r & 31)) & 1u) old_of[remap[r]] = (u32)r;
}

struct GcTarget {      // the fresh persistent arrays the live rows are gathered into
	double *coord;
	u32 *row_slot, *root, *live, *ideal, *inc_off, *inc_len, *adj_off, *adj_len, *inc_pool, *adj_pool;
};
// one thread per NEW row (blockDim multiple of 32 so a warp owns one word of the bitsets)
__global__ void __launch_bounds__(K_THREADS) k_gc_gather(DevState S, GcTarget T, u32 n_live, const u32 *remap,
                                                         const u32 *old_of, const u32 *new_inc_off, const u32 *new_adj_off)
{
	const size_t cap = S.cap_rows;
	const u32 npad = (n_live + 31) & ~31u;
	B200_GRID_STRIDE(n, npad) {
		const bool in = n < n_live;
		const u32 o = in ? old_of[n] : 0;
		const u32 idl = __ballot_sync(0xffffffffu, in && ((S.ideal[o >> 5] >> (o & 31)) & 1u));
		const u32 liv = __ballot_sync(0xffffffffu, in);
		if ((threadIdx.x & 31) == 0) { T.ideal[n >> 5] = idl; T.live[n >> 5] = liv; }
		if (!in) continue;
		for (int j = 0; j < S.d; j++) T.coord[j * cap + n] = S.coord[j * cap + o];
		T.row_slot[n] = S.row_slot[o];
		T.root[n] = S.root[o];
		const u32 il = S.inc_len[o], io = S.inc_off[o], ni = new_inc_off[n];
		T.inc_off[n] = ni;
		T.inc_len[n] = il;
		for (u32 q = 0; q < il; q++) T.inc_pool[ni + q] = S.inc_pool[io + q];
		const u32 al = S.adj_len[o], ao = S.adj_off[o], na = new_adj_off[n];
		T.adj_off[n] = na;
		T.adj_len[n] = al;
		for (u32 q = 0; q < al; q++) T.adj_pool[na + q] = remap[S.adj_pool[ao + q]];
	}
}
__global__ void k_gc_finish(DevState S, u32 n_live, const u32 *inc_total, const u32 *adj_total)
{
	if (threadIdx.x || blockIdx.x) return;
	S.ctl->nrows = n_live;
	S.ctl->n_live = n_live;
	S.ctl->inc_used = *inc_total;
	S.ctl->adj_used = *adj_total;
}

// ====================================================================================================
// Small-cut path: a streaming K1 with no block-level synchronisation followed by single-CTA "tail"
// kernels that run every remaining stage of the cut as __syncthreads()-separated phases.  A cut of a
// 10^6-vertex polytope touches a few hundred vertices; as separate launches those stages cost ~20
// launch latencies, as phases of one CTA they cost ~20 block barriers.
// ====================================================================================================

// K1, streaming form: all loads of a tile are independent (no liveness test before the coordinate
// loads); the rare non-PLUS rows get their class byte written and are appended to one unordered
// list through an atomic each -- no shared memory, no barrier.  The tail sorts the list.
template <int D, bool FROMDEV, int ITREQ>
__global__ void __launch_bounds__(K_THREADS, (ITREQ >= 4 ? 2 : ITREQ == 2 ? 4 : 8)) k_classify_lists(DevState S, CutParams Parg, const double *vals,
                                                              const unsigned char *ideal, u64 vi, u32 nrows_host, u32 tile_lo, u32 tile_hi)
{
	const int d = D > 0 ? D : S.d;
	double h[D > 0 ? D : B200_MAXD];
	double alpha;
	if (FROMDEV) {
		for (int j = 0; j < d; j++) h[j] = vals[vi * d + j];
		alpha = (ideal && ideal[vi]) ? 0.0 : -1.0;
	} else {
		for (int j = 0; j < d; j++) h[j] = Parg.h[j];
		alpha = Parg.alpha;
	}
	const double hi0 = __dadd_rn(alpha, POLY_EPS_D), mid0 = __dadd_rn(alpha, 1.0e-2 * POLY_EPS_D), lo0 = __dsub_rn(alpha, POLY_EPS_D);
	const double hi1 = POLY_EPS_D, mid1 = 1.0e-2 * POLY_EPS_D, lo1 = -POLY_EPS_D;
	if (blockIdx.x == 0 && threadIdx.x == 0) {          // publish the halfspace for the later stages
		CutParams P = Parg;
		if (FROMDEV) {
			double hh = 0;
			for (int j = 0; j < B200_MAXD; j++) { P.h[j] = j < d ? h[j] : 0.0; hh = __dadd_rn(hh, __dmul_rn(P.h[j], P.h[j])); }
			P.alpha = alpha;
			P.hh = hh;
		}
		P.hi[0] = hi0; P.mid[0] = mid0; P.lo[0] = lo0;
		P.hi[1] = hi1; P.mid[1] = mid1; P.lo[1] = lo1;
		*S.cur = P;
		S.facet_cnt[P.facet] = 0;
		S.facet_alive[P.facet] = 1;
	}
	// rows [tile_lo, tile_hi) * 2048: the whole polytope on one GPU, this rank's share when sharded
	(void)nrows_host;
	const size_t cap = S.cap_rows;
	constexpr int NIT = B200_TILE / (2 * K_THREADS);
	constexpr int IT = (D > 0 && D <= 8) ? (ITREQ < NIT ? ITREQ : NIT) : 1;   // loads kept in flight per thread: IT * D double2
	for (u32 tg = blockIdx.x; tg < (tile_hi - tile_lo) * (NIT / IT); tg += gridDim.x) {
		const u32 tile = tile_lo + tg / (NIT / IT), it0 = (tg % (NIT / IT)) * IT;
		u32 lw[IT], iw[IT];
		double2 x[IT][D > 0 ? D : B200_MAXD];
#pragma unroll
		for (int it = 0; it < IT; it++) {
			const u32 r = tile * B200_TILE + (it0 + it) * 2 * K_THREADS + 2 * threadIdx.x;
			lw[it] = S.live[r >> 5] >> (r & 31);
			iw[it] = S.ideal[r >> 5] >> (r & 31);
#pragma unroll
			for (int j = 0; j < d; j++) x[it][j] = *reinterpret_cast<const double2 *>(S.coord + j * cap + r);
		}
#pragma unroll
		for (int it = 0; it < IT; it++) {
			const u32 r = tile * B200_TILE + (it0 + it) * 2 * K_THREADS + 2 * threadIdx.x;
			double t0 = __dmul_rn(h[0], x[it][0].x), t1 = __dmul_rn(h[0], x[it][0].y);
#pragma unroll
			for (int j = 1; j < d; j++) {
				t0 = __dadd_rn(t0, __dmul_rn(h[j], x[it][j].x));
				t1 = __dadd_rn(t1, __dmul_rn(h[j], x[it][j].y));
			}
			u8 c[2] = {CLS_DEAD, CLS_DEAD};
			const double t[2] = {t0, t1};
#pragma unroll
			for (int s = 0; s < 2; s++) {
				if (!((lw[it] >> s) & 1u)) continue;
				const bool id = (iw[it] >> s) & 1u;
				const double hi = id ? hi1 : hi0, mid = id ? mid1 : mid0, lo = id ? lo1 : lo0;
				c[s] = t[s] > hi ? CLS_PLUS : t[s] > mid ? CLS_ZP : t[s] > lo ? CLS_ZERO : CLS_MINUS;
				if (c[s] != CLS_PLUS) {                  // rare; PLUS rows already read PLUS (invariant between cuts)
					S.cls[r + s] = c[s];
					const u32 pos = atomicAdd(&S.ctl->n_list, 1u);
					if (pos < B200_VIS_MAX) S.nplist[pos] = r + s;
					if (c[s] == CLS_ZP) atomicAdd(&S.ctl->n_zp, 1u);
					if (t[s] < lo) { atomicAdd(&S.ctl->n_strict, 1u); atomicMin(&S.ctl->min_strict_row, r + s); }
				}
			}
		}
	}
}

#define TAIL_THREADS 1024
#define TAIL_CTAS 8            // medium cuts: one thread-block cluster (portable size), phases separated by cluster barriers;
                               // tiny cuts: a single CTA, phases separated by __syncthreads()
template <int NC> __device__ __forceinline__ u32 tail_rank()
{
	if (NC == 1) return 0;
	u32 r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
template <int NC> __device__ __forceinline__ void tail_sync()
{
	if (NC == 1) __syncthreads();
	else asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#define TAIL_LOOP(i, n) for (u32 i = ctid; i < (u32)(n); i += NC * TAIL_THREADS)
#define TAIL_SYNC() tail_sync<NC>()
// Items whose work is gathers (row -> lists -> neighbours) are spread over ALL warps of the cluster, a few lanes
// per warp: measured on B200 a warp-level memory instruction costs ~4 cycles per distinct sector it touches, per
// warp, so a chain of gathers finishes 32/c times sooner in a warp with c active lanes -- and the other warps
// would sit at the barrier anyway.  Item w*c + l goes to lane l < c of warp w (consecutive items stay together).
#define TAIL_SPREAD(i, n)                                                                                                  \
	for (u32 _b = 0, _n = (u32)(n); _b < _n; _b += NC * TAIL_THREADS)                                                      \
		for (u32 _m = min(_n - _b, (u32)(NC * TAIL_THREADS)), _c = (_m + NC * (TAIL_THREADS / 32) - 1) / (NC * (TAIL_THREADS / 32)), \
		         _l = threadIdx.x & 31, i = _b + (rank * (TAIL_THREADS / 32) + (threadIdx.x >> 5)) * _c + _l, _once = 1;     \
		     _once && _l < _c && i < _n; _once = 0)
#define TP(k) do { if (ctid == 0) S.dbg[k] = b200_globaltimer(); } while (0)

// The staged record lives in mapped pinned host memory.  Every thread that wrote payload has fenced at
// system scope (and a barrier lies in between); here the header words are written, fenced, and only
// then the sequence number the host is spinning on.  Called by all threads of CTA 0; its first warp works.
__device__ __forceinline__ void tail_stage_header(const DevState &S, u32 extra_status, bool header_only)
{
	CutCtl *c = S.ctl;
	const StageLayout L = stage_layout(*c, S.d);
	const bool fits = header_only || L.total <= S.cap_stage;
	if (threadIdx.x < 32) {
		if (threadIdx.x < sizeof(CutCtl) / 4) {
			u32 v = ((const u32 *)c)[threadIdx.x];
			if (threadIdx.x == offsetof(CutCtl, status) / 4) v |= extra_status | (fits ? 0u : (u32)ST_OVF_STAGE);
			if (threadIdx.x == offsetof(CutCtl, stage_bytes) / 4) v = (u32)L.total;
			((volatile u32 *)S.stage)[threadIdx.x] = v;
		}
		__threadfence_system();
		__syncwarp();
		if (threadIdx.x == 0) *(volatile u32 *)(S.stage + B200_STAGE_SEQ) = S.cur->seq;
	}
}
__device__ __forceinline__ void tail_reset_for_next_cut(const DevState &S)
{
	// the counters the streaming K1 accumulates into (one thread, after the header was staged)
	S.ctl->n_strict = 0;
	S.ctl->min_strict_row = B200_NONE;
	S.ctl->n_zp = 0;
	S.ctl->n_list = 0;
}

// phases after K4: adjacency build, commit, delta record (also the body of k_tail2)
template <int NC> __device__ void tail_adjacency_and_pack(const DevState &S, u32 *ws, bool header_only, bool wrote_payload)
{
	CutCtl *c = S.ctl;
	const u32 rank = tail_rank<NC>(), ctid = rank * TAIL_THREADS + threadIdx.x;
	// Nothing in here writes the status word or the persistent counters before the last barrier (capacity
	// overflows found by CTA 0 travel in scratch_flag, the commit waits), so no barrier is needed on entry and
	// the bodies may read nrows / adj_used while other warps are still working.
	const u32 st_in = c->status;
	const u32 OVF_P = 8u, OVF_A = 16u;
	TP(20);
	bool go = !(st_in & ST_SKIP_B);
	if (go) {
		if (rank == 0) {
			if (c->n_pairs > S.cap_pairs || c->n_surv > S.cap_pairs) {
				if (threadIdx.x == 0) atomicOr(&c->scratch_flag, OVF_P);
			} else {
				const u32 n = c->n_new;
				u32 carry = 0;
				for (u32 base = 0; base < n; base += TAIL_THREADS) {
					u32 j = base + threadIdx.x, v = j < n ? S.new_padj_len[j] + S.deg[j] : 0, tot;
					u32 e = block_excl_scan(v, ws, tot);
					if (j < n) S.adj_base[j] = carry + e;
					carry += tot;
				}
				if (threadIdx.x == 0) {
					c->adj_new = carry;
					if ((u64)c->adj_used + carry > S.cap_adj) atomicOr(&c->scratch_flag, OVF_A);
				}
			}
		}
		TAIL_SYNC();
		TP(21);
		go = !(c->scratch_flag & (OVF_P | OVF_A));
		if (go) {
			// adjacency offsets are known: place the PLUS neighbours
			TAIL_SPREAD(j, c->n_new) adj_place(S, j);
			TP(22);
			TAIL_SPREAD(p, c->n_pairs) adj_pair_fill(S, p);   // independent of the placement: offsets are derived, not read back
			TAIL_SYNC();
			TP(23);
			TAIL_SPREAD(j, c->n_new) adj_sort(S, j);
		}
	}
	// single-launch cut: the record went to host memory from this grid, every writer orders its payload before
	// the header / sequence number (late, so that the PCIe writes have drained by now); with k_tail2 as a
	// separate launch the grid boundary has done that already
	TP(24);
	if (wrote_payload) __threadfence_system();
	TAIL_SYNC();
	TP(25);
	if (rank == 0) {
		if (threadIdx.x == 0) {
			const u32 fl = c->scratch_flag;
			if (fl & OVF_P) c->status |= ST_OVF_PAIRS;
			if (fl & OVF_A) c->status |= ST_OVF_ADJ;
			if (go) {                                  // commit: every reader of these counters is past the barrier
				c->n_live = c->n_live + c->n_new - (c->n_minus + c->n_zero);
				c->nrows += c->n_new;
				c->slot_cnt += c->n_new;
				c->inc_used += c->inc_new;
				c->adj_used += c->adj_new;
			}
		}
		__syncthreads();
		tail_stage_header(S, 0, header_only);
		__syncthreads();
		if (threadIdx.x == 0) tail_reset_for_next_cut(S);
	}
}

// mode 0: the whole rest of the cut; mode 1: stop after building K4's bit matrix (the multi-block
// k4_filter / k4_contain and k_tail2 follow)
#define TAIL_MODE_STOP_AT_K4 1    // stop after building K4's matrices: k4_filter, k4_contain, k_tail2 follow
#define TAIL_MODE_FUSED_K1 2      // small polytope: this (single) CTA classifies the rows itself, no K1 launch
template <int NC> __global__ void __launch_bounds__(TAIL_THREADS, 1) k_tail(DevState S, int mode_bits, int header_only, CutParams Parg, u32 nrows_host)
{
	const int mode = mode_bits & TAIL_MODE_STOP_AT_K4;
	__shared__ u32 ws[99];
	__shared__ u32 slist[B200_VIS_MAX];
	__shared__ u32 soff[B200_VIS_MAX + 1];
	const u32 rank = tail_rank<NC>(), ctid = rank * TAIL_THREADS + threadIdx.x;
	cudaGridDependencySynchronize();      // K1 (and the exchange kernels) precede this launch
	CutCtl *c = S.ctl;
	TP(0);
	if (NC == 1 && (mode_bits & TAIL_MODE_FUSED_K1)) {
		// K1 inside the tail: a polytope of a few thousand rows (the Benson-loop regime) is classified
		// by this CTA in a few hundred cycles; one launch per cut instead of two
		if (threadIdx.x == 0) {
			*S.cur = Parg;
			S.facet_cnt[Parg.facet] = 0;
			S.facet_alive[Parg.facet] = 1;
		}
		for (u32 r = threadIdx.x; r < nrows_host; r += TAIL_THREADS) {
			bool strict, zp;
			const u8 cl = classify_row(S, Parg, r, strict, zp);
			if (cl == CLS_DEAD || cl == CLS_PLUS) continue;
			S.cls[r] = cl;
			const u32 pos = atomicAdd(&c->n_list, 1u);
			if (pos < B200_VIS_MAX) S.nplist[pos] = r;
			if (zp) atomicAdd(&c->n_zp, 1u);
			if (strict) { atomicAdd(&c->n_strict, 1u); atomicMin(&c->min_strict_row, r); }
		}
		__syncthreads();
	}
	// ---- P0: reset per-cut outputs, decide, sort K1's unordered list into the ascending visited list
	const u32 n_list = c->n_list;
	if (ctid == 0) {
		c->status = 0;
		c->min_strict_slot = B200_NONE;
		c->n_zp_projected = 0;
		c->n_new = c->inc_new = c->padj_new = 0;
		c->n_minus = c->n_zero = 0;
		c->n_pairs = c->adj_new = c->n_dead_facets = 0;
		c->n_live_scanned = c->n_live;
		c->n_local = c->wl = c->mpad = c->n_surv = 0;
		c->scratch_flag = 0;
		c->n_vis = n_list <= B200_VIS_MAX ? n_list : 0;
		if (c->n_strict == 0) {                          // nothing to cut: redundant (bslv_poly.c:132-136)
			c->status |= ST_REDUNDANT;
			S.facet_alive[S.cur->facet] = 0;
			if (n_list > B200_VIS_MAX) c->status |= ST_NEED_BIG;   // cannot even undo the class marks from the list
		} else {
			c->min_strict_slot = S.row_slot[c->min_strict_row];
			if (n_list > B200_VIS_MAX) c->status |= ST_NEED_BIG;
		}
	}
	if (n_list <= B200_VIS_MAX) {
		// on-plane vertices (rare): clear the shared-facet masks the half-edge evaluation ORs into
		if (n_list > c->n_strict) TAIL_LOOP(x, n_list * (B200_MAXINC / 64)) S.zmask[x] = 0;
		// every CTA stages the list in shared memory; an element's position in the sorted order is
		// the number of smaller elements (rows are distinct)
		for (u32 x = threadIdx.x; x < n_list; x += TAIL_THREADS) slist[x] = S.nplist[x];
		__syncthreads();
		// CTA `rank` ranks its slice of the elements; 8 consecutive lanes share one element and count
		// over interleaved eighths of the list (n^2 / (8192 threads) comparisons each)
		const u32 per = (n_list + NC - 1) / NC, e0 = rank * per, e1 = min(n_list, e0 + per);
		const u32 part = threadIdx.x & 7;
		for (u32 eb = e0; eb < e1; eb += TAIL_THREADS / 8) {      // block-uniform trip count
			const u32 i = eb + (threadIdx.x >> 3);
			const u32 key = i < e1 ? slist[i] : 0;
			u32 rk = 0;
			for (u32 x = part; x < n_list; x += 8) rk += (slist[x] < key);
			rk += __shfl_xor_sync(0xffffffffu, rk, 1);
			rk += __shfl_xor_sync(0xffffffffu, rk, 2);
			rk += __shfl_xor_sync(0xffffffffu, rk, 4);
			if (part == 0 && i < e1) S.vis[rk] = key;
		}
	}
	TAIL_SYNC();
	if (c->status & (ST_REDUNDANT | ST_NEED_BIG)) {
		if (!(c->status & ST_NEED_BIG)) TAIL_SPREAD(i, c->n_vis) reset_class(S, i);
		if (rank == 0) tail_stage_header(S, 0, header_only);
		TAIL_SYNC();
		if (ctid == 0) tail_reset_for_next_cut(S);
		return;
	}
	const CutParams &P = *S.cur;
	const u32 n_vis = c->n_vis;
	TP(1);
	// ---- P1: ZERO+ closure (rare)
	if (c->n_zp) {
		for (;;) {
			bool any = false;
			TAIL_SPREAD(i, n_vis) any |= zp_activate(S, P, i);
			if (any) atomicOr(&c->scratch_flag, 1u);
			TAIL_SYNC();
			const u32 f = c->scratch_flag;
			TAIL_SYNC();
			if (!f) break;
			if (ctid == 0) c->scratch_flag = 0;
			TAIL_SYNC();
		}
	}
	// ---- P2: half-edge offsets.  Every CTA derives them itself into shared memory (the visited list is short),
	// so no cluster barrier separates the scan from the evaluation; CTA 0 also stores them for the later phases
	for (u32 x = threadIdx.x; x < n_vis; x += TAIL_THREADS) slist[x] = S.vis[x];      // read back by the same thread below
	u32 H = 0;
	for (u32 base = 0; base < n_vis; base += TAIL_THREADS) {
		u32 i = base + threadIdx.x, v = 0, tot;
		if (i < n_vis) {
			const u32 r = slist[i];
			v = is_visited_class(S.cls[r]) ? S.adj_len[r] : 0;
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
	if (H > S.cap_he) {                    // every CTA sees the same total and takes this branch
		// The verdict goes into the control block itself (after a barrier: other CTAs may still be reading the status
		// word above), not only into the staged copy: when the host has already queued k4_filter / k4_contain /
		// k_tail2 behind this launch they must see it and do nothing.
		TAIL_SYNC();
		if (ctid == 0) c->status |= ST_NEED_BIG;
		TAIL_SYNC();
		if (rank == 0) tail_stage_header(S, 0, header_only);
		TAIL_SYNC();
		if (ctid == 0) tail_reset_for_next_cut(S);
		return;
	}
	TP(2);
	// ---- P3: evaluate every (visited vertex, neighbour) pair; the owner of a half-edge is found by bisection
	TAIL_SPREAD(e, H) {
		u32 lo = 0, hi = n_vis;            // soff[lo] <= e < soff[hi]
		while (hi - lo > 1) {
			const u32 mid = (lo + hi) >> 1;
			if (soff[mid] <= e) lo = mid;
			else hi = mid;
		}
		S.he_own[e] = lo;
		he_eval_at(S, e, lo, slist[lo], soff[lo]);
	}
	TAIL_SYNC();
	TP(3);
	// ---- P4a: sizes per visited vertex, spread over the whole cluster (the strided list accesses are LSU work)
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
		if ((threadIdx.x & 31) == 0) {               // one atomic per warp instead of one per visited vertex
			if (nm) atomicAdd(&c->n_minus, nm);
			if (nz) atomicAdd(&c->n_zero, nz);
		}
		if (bad) atomicOr(&c->status, (u32)ST_ERR_DEGENERATE);
	}
	TAIL_SYNC();
	TP(16);
	// ---- P4b: offsets and capacity plan.  Every CTA scans the (contiguous) counts itself, so no further cluster
	// barrier separates the plan from the emission; CTA 0 alone updates the control block
	{
		u32 carry[3] = {0, 0, 0};
		const bool bad = (c->status & ST_ERR_DEGENERATE) != 0;      // (only the OVF bits can change under this read)
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
		u32 st = 0;
		if ((u64)c->nrows + carry[0] > S.cap_rows) st |= ST_OVF_ROWS;
		if ((u64)c->inc_used + carry[1] > S.cap_inc) st |= ST_OVF_INC;
		if (carry[2] > S.cap_padj) st |= ST_OVF_PADJ;
		if (ctid == 0) {
			c->n_new = carry[0];
			c->inc_new = carry[1];
			c->padj_new = carry[2];
			S.facet_cnt[S.cur->facet] = carry[0];      // every new row lies on the new facet
			if (st) atomicOr(&c->status, st);
		}
		// block-wide (and, the inputs being the same, cluster-wide) agreement on whether the cut can proceed;
		// the barrier also orders this CTA's base3 stores before its own reads below
		if (__syncthreads_or(st != 0 || bad)) {
			if (rank == 0) tail_stage_header(S, 0, header_only);
			TAIL_SYNC();
			if (ctid == 0) tail_reset_for_next_cut(S);
			return;
		}
	}
	TP(4);
	// ---- P5: new rows + rewiring (per half-edge) and copies + retirement (per vertex): independent
	TAIL_SPREAD(x, 2 * H) {                            // geometry and incidence of a new row go to different warps
		const int part = x >= H;
		he_emit_part(S, P, part ? x - H : x, part);
	}
	TP(17);
	TAIL_SPREAD(i, n_vis) he_finish_vertex(S, P, i);
	TP(18);
	TAIL_SYNC();
	TP(5);
	// ---- P6: dead facets (needs the final facet counts) ‖ K4's matrix shape and the clearing of its column matrix
	// (the columns were assigned as the rows were emitted, so their number is final since the barrier above)
	const u32 M = c->n_new;
	// every thread derives the matrix shape itself (no barrier between plan and the stores below)
	const u32 wl = (c->n_local + 63) / 64, mpad = (M + 63) & ~63u;
	const bool bits_ovf = k4_words(wl, mpad, c->n_local) > S.cap_bits;
	if (ctid == 0) k4_plan(S);
	if (!bits_ovf) {
		u64 *tb = k4_tbits(S, wl, mpad);
		for (u64 x = ctid; x < (u64)c->n_local * (mpad / 64); x += NC * TAIL_THREADS) tb[x] = 0;
	}
	TAIL_SPREAD(i, n_vis) collect_dead_facets(S, i);
	TP(19);
	TAIL_SYNC();
	TP(6);
	// ---- P7: delta record ‖ K4 bit matrices.  Everything the record holds is final now (new rows, parents, retired
	// slots, dead facets); the stores go to mapped host memory and drain over PCIe while the pair test and the
	// adjacency build run
	bool wrote_payload = false;
	if (!header_only && !bits_ovf) {          // (on a bit-matrix overflow the host grows and re-runs; the record is packed then)
		const StageLayout L = stage_layout(*c, S.d);
		if (L.total <= S.cap_stage) {
			const u64 n = (u64)M * S.d + M + n_vis + c->n_dead_facets;
			const u32 first_row = c->nrows;             // not committed yet
			for (u64 e = ctid; e < n; e += NC * TAIL_THREADS) pack_delta_item(S, L, e, first_row);
			wrote_payload = ctid < n;
		}
	}
	if (!bits_ovf) {
		{
			const u32 nrows = c->nrows, f = P.facet;
			TAIL_SPREAD(j, M) k4_build_row_at(S, j, nrows, f, wl, mpad);
		}
		TAIL_SYNC();
		TP(7);
		if (mode == 1) return;                            // k4_filter, k4_contain, k_tail2 follow
		// the pair filter is M^2 integer work: 8 SMs only pay off for small M, beyond that the
		// multi-block k4_filter / k4_contain (all 148 SMs) follow
		if (M > (NC == 1 ? B200_K4_SMALL / 2 : B200_K4_SMALL)) {
			if (rank == 0) tail_stage_header(S, ST_K4_PENDING, header_only);
			return;
		}
		// ---- P7: pair test inside the cluster
		for (u32 p = ctid; p < M * M; p += NC * TAIL_THREADS) {
			const u32 a = p / M, b = p % M;
			if (a < b) k4_filter_pair_in(S, S.bits, wl, mpad, a, b, k4_threshold(S, true));
		}
		TAIL_SYNC();
		TP(8);
		if (c->n_surv <= S.cap_pairs) {
			const u32 ns = c->n_surv;
			for (u32 s = ctid >> 5; s < ns; s += NC * TAIL_THREADS / 32) k4_contain_warp(S, S.bits, s, threadIdx.x & 31, M, wl, mpad);
		}
		TAIL_SYNC();
	} else {
		TAIL_SYNC();
		if (mode == 1) return;
	}
	TP(9);
	// ---- P8: adjacency, commit, delta record
	tail_adjacency_and_pack<NC>(S, ws, header_only, wrote_payload);
	TP(10);
}

template <int NC> __global__ void __launch_bounds__(TAIL_THREADS, 1) k_tail2(DevState S, int header_only)
{
	__shared__ u32 ws[33];
	const u32 ctid = tail_rank<NC>() * TAIL_THREADS + threadIdx.x;
	cudaGridDependencySynchronize();
	TP(11);
	tail_adjacency_and_pack<NC>(S, ws, header_only, false);
	TP(12);
}

__global__ void k_reset_small(DevState S)
{
	if (threadIdx.x || blockIdx.x) return;
	S.ctl->n_strict = 0;
	S.ctl->min_strict_row = B200_NONE;
	S.ctl->n_zp = 0;
	S.ctl->n_list = 0;
}

// L2 flush for measurements: a read-only sweep over a buffer larger than L2 leaves clean lines behind
// (a memset would leave dirty ones whose write-back then competes with the timed kernel's reads)
__global__ void __launch_bounds__(K_THREADS) k_flush_read(const uint4 *buf, size_t n, unsigned *sink)
{
	unsigned acc = 0;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint4 v = buf[i];
		acc ^= v.x ^ v.y ^ v.z ^ v.w;
	}
	if (acc == 0x9e3779b9u) *sink = acc;
}

// ------------------------------------------------------------------ multi-GPU exchange
// State is replicated, K1 is sharded by row range.  After its share of K1 a rank packs what it
// found -- trigger counters and its non-PLUS rows with their classes, ascending -- into a fixed-size
// record; one all-gather later every rank merges the records (rank ranges are ascending, so the
// concatenation is the ordered visited list) and runs the rest of the cut identically.
__global__ void __launch_bounds__(TAIL_THREADS, 1) k_xchg_pack(DevState S)
{
	CutCtl *c = S.ctl;
	u32 *out = S.xchg_send;
	const u32 n = c->n_list, over = n > B200_XCHG_CAP;
	for (u32 i = threadIdx.x; i < n && !over; i += TAIL_THREADS) {
		const u32 row = S.nplist[i];
		out[4 + i] = row | ((u32)S.cls[row] << 30);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		out[0] = c->n_strict;
		out[1] = c->min_strict_row;
		out[2] = c->n_zp;
		out[3] = over ? B200_NONE : n;
		c->n_strict = 0;                 // local accumulators: the merged values are written by k_xchg_merge
		c->min_strict_row = B200_NONE;
		c->n_zp = 0;
		c->n_list = 0;
	}
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) k_xchg_merge(DevState S, u32 nranks)
{
	__shared__ u32 off[66];
	CutCtl *c = S.ctl;
	if (threadIdx.x == 0) {
		u32 ns = 0, mr = B200_NONE, nz = 0, tot = 0, over = 0;
		for (u32 g = 0; g < nranks; g++) {
			const u32 *r = S.xchg_recv + (size_t)g * B200_XCHG_WORDS;
			ns += r[0];
			mr = min(mr, r[1]);
			nz += r[2];
			off[g] = tot;
			if (r[3] == B200_NONE) over = 1; else tot += r[3];
		}
		off[nranks] = tot;
		if (tot > B200_VIS_MAX) over = 1;
		c->n_strict = ns;
		c->min_strict_row = mr;
		c->n_zp = nz;
		c->n_list = over ? B200_VIS_MAX + 1 : tot;     // overflow: the multi-kernel path re-classifies everything unsharded
		off[65] = over;
	}
	__syncthreads();
	if (off[65]) return;
	for (u32 g = 0; g < nranks; g++) {
		const u32 *r = S.xchg_recv + (size_t)g * B200_XCHG_WORDS + 4;
		const u32 n = off[g + 1] - off[g];
		for (u32 i = threadIdx.x; i < n; i += TAIL_THREADS) {
			const u32 e = r[i], row = e & 0x3FFFFFFFu;
			S.nplist[off[g] + i] = row;
			S.cls[row] = (u8)(e >> 30);
		}
	}
}

// ------------------------------------------------------------------ K6: dual all-pairs adjacency
__global__ void __launch_bounds__(K_THREADS) k6_build(DevState S, u32 mpad)
{
	B200_GRID_STRIDE(r, S.ctl->nrows) k6_set_row_bits(S, (u32)r, mpad);
}
__global__ void k6_begin(DevState S, u32 M, u32 wl, u32 mpad)
{
	if (threadIdx.x || blockIdx.x) return;
	CutCtl *c = S.ctl;
	c->status = 0;
	c->n_new = M;
	c->wl = wl;
	c->mpad = mpad;
	c->n_local = 0;
	c->n_surv = c->n_pairs = 0;
}

// multi-kernel path: the record was written by a whole grid (k_pack_delta); the launch boundary orders
// it before this single store of the sequence number the host spins on
__global__ void k_publish(DevState S, u32 seq)
{
	if (threadIdx.x || blockIdx.x) return;
	__threadfence_system();
	*(volatile u32 *)(S.stage + B200_STAGE_SEQ) = seq;
}
// re-launches inside one cut (pending K4, re-run after growth) publish under a fresh number
__global__ void k_set_seq(DevState S, u32 seq)
{
	if (threadIdx.x || blockIdx.x) return;
	S.cur->seq = seq;
}
