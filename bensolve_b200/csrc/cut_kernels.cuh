// sm_100a kernels of the cut (K1 classify, K2 compact, K3 edge-intersect, K4 pair adjacency,
// K5 rewire/append).  All launches of one cut are stream-ordered; every kernel sizes its work
// from the device-side CutCtl, so the host never has to read a count back between stages.
// Tensor cores are deliberately not used: the path is matrix-vector and list/bitset work.
#pragma once
#include <cuda_runtime.h>

#include "cut_bodies.h"

#define K_THREADS 256
#define SCAN_THREADS 1024
#define ST_SKIP_A (ST_REDUNDANT | ST_OVF_A | ST_ERR_DEGENERATE)
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
	S.facet_cnt[P.facet] = 0;
	S.facet_alive[P.facet] = 1;
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
	if (S.ctl->status & ST_SKIP_A) return;
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
	if (S.ctl->status & ST_SKIP_A) return;
	B200_GRID_STRIDE(i, S.ctl->n_vis) collect_dead_facets(S, (u32)i);
}

// ------------------------------------------------------------------ K4 (list form) and adjacency build
__global__ void __launch_bounds__(K_THREADS) k_pairs_reset(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		S.ctl->n_pairs = 0;
		S.ctl->status &= ~(u32)ST_OVF_B;
	}
	B200_GRID_STRIDE(j, S.ctl->n_new) S.deg[j] = 0;
}

__global__ void __launch_bounds__(K_THREADS) k_pairs(DevState S)
{
	if (S.ctl->status & ST_SKIP_A) return;
	const u32 M = S.ctl->n_new;
	B200_GRID_STRIDE(p, (u64)M * M) pair_test(S, p, M);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_adj_scan(DevState S)
{
	__shared__ u32 ws[33];
	CutCtl *c = S.ctl;
	if (c->status & ST_SKIP_A) return;
	if (c->n_pairs > S.cap_pairs) {
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
