// Wave path, per-item logic shared by the sm_100a kernels (wave_kernels.cuh) and the host-side test double
// (-DB200_EMULATE): look-ahead classification of every row against the pending halfspaces, formation of a wave
// of commuting cuts, the bases each cut of a wave appends at, the commit.
//
// What is replaced is still bslv_poly.c:104-151 (poly__add_vrtx: scan :121-128, poly__cut :562-709, pair loop
// :138-143); the reference does one halfspace at a time, so the look-ahead and the waves have no counterpart
// there -- the results are the reference's because cuts of one wave commute.
#pragma once
#include "cut_bodies.h"

#if defined(__CUDA_ARCH__)
#define B200_ATOMIC_MAX(p, v) atomicMax((p), (v))
#else
static inline u32 b200_fetch_max(u32 *p, u32 v) { u32 o = *p; if (v > o) *p = v; return o; }
#define B200_ATOMIC_MAX(p, v) b200_fetch_max((p), (v))
#endif

B200_HD double wv_abs(double x) { return x < 0 ? -x : x; }

// The shared state seen by one cut of a wave: control block and halfspace of `slot`, scratch of wave position `wpos`.
B200_HD DevState wave_view(const DevState &S, const WaveDev &W, u32 slot, u32 wpos)
{
	DevState V = S;
	const size_t L = B200_WAVE_LIST;
	V.ctl = W.ctl + slot;
	V.cur = W.cur + slot;
	V.nplist = W.list + (size_t)slot * L;
	V.vis = W.vis + (size_t)wpos * L;
	V.cnt3 = W.cnt3 + (size_t)wpos * 3 * L;
	V.base3 = W.base3 + (size_t)wpos * 3 * L;
	V.dead_slots = W.dead_slots + (size_t)wpos * L;
	V.he_off = W.he_off + (size_t)wpos * (L + 1);
	V.he_own = W.he_own + (size_t)wpos * W.cap_he;
	V.he_inc = W.he_inc + (size_t)wpos * W.cap_he;
	V.he_k = W.he_k + (size_t)wpos * W.cap_he;
	V.he_rank = W.he_rank + (size_t)wpos * W.cap_he;
	V.he_incpre = W.he_incpre + (size_t)wpos * W.cap_he;
	V.he_flag = W.he_flag + (size_t)wpos * W.cap_he;
	V.zmask = W.zmask + (size_t)wpos * L * (B200_MAXINC / 64);
	V.zlong = nullptr;                 // (cuts with such vertices leave the wave: ST_NEED_BIG in k_wave_tailA)
	V.zlong_words = 0;
	V.padj = W.padj + (size_t)wpos * W.cap_new;
	V.new_padj_off = W.new_padj_off + (size_t)wpos * W.cap_new;
	V.new_padj_len = W.new_padj_len + (size_t)wpos * W.cap_new;
	V.new_parent = W.new_parent + (size_t)wpos * W.cap_new;
	V.deg = W.deg + (size_t)wpos * W.cap_new;
	V.adj_fill = W.adj_fill + (size_t)wpos * W.cap_new;
	V.adj_base = W.adj_base + (size_t)wpos * W.cap_new;
	V.pair_a = W.pair_a + (size_t)wpos * W.cap_pairs;
	V.pair_b = W.pair_b + (size_t)wpos * W.cap_pairs;
	V.surv_a = W.surv_a + (size_t)wpos * W.cap_pairs;
	V.surv_b = W.surv_b + (size_t)wpos * W.cap_pairs;
	V.facet_epoch = W.facet_epoch + (size_t)wpos * W.cap_facets;
	V.facet_local = W.facet_local + (size_t)wpos * W.cap_facets;
	V.dead_facets = W.dead_facets + (size_t)wpos * W.cap_facets;
	V.bits = W.bits + (size_t)wpos * W.cap_bits;
	V.cap_padj = W.cap_new;
	V.cap_pairs = W.cap_pairs;
	V.cap_bits = W.cap_bits;
	V.cap_he = W.cap_he;
	return V;
}

// ---------------------------------------------------------------- look-ahead classification
// Code of a row for a list: 0xFF = safely PLUS (not listed).  A row is safely PLUS when it clears the PLUS
// threshold (bslv_poly.c:596) by a guard band far wider than any rounding error of a vertex later created on an
// edge between two such rows, so such a vertex is PLUS too and a cut never has to look at it.
B200_HD u32 wave_code(double t, int id, const CutParams &P, double xinf)
{
	const double thr = id ? 0.0 : P.alpha;
	const double g = B200_WV_GUARD * (wv_abs(thr) + P.h1 * xinf);
	if (t > P.hi[id] + g) return 0xFFu;
	u32 c = class_of(t, id, P);
	if (t < P.lo[id]) c |= B200_WV_STRICT;
	return c;
}
B200_HD void wave_list_append(const WaveDev &W, u32 slot, u32 row, u32 code)
{
	const u32 pos = B200_ATOMIC_ADD(&W.ctl[slot].n_list, 1u);
	if (pos < B200_WAVE_LIST) W.list[(size_t)slot * B200_WAVE_LIST + pos] = row | (code << B200_WV_ROW_BITS);
}
// (sharded pass) the entry goes to this rank's exchange record instead
B200_HD void wave_send_append(const WaveDev &W, u32 slot, u32 row, u32 code)
{
	const u32 pos = B200_ATOMIC_ADD((u32 *)W.xsend, 1u);        // low word of xsend[0] (little endian): the entry count
	if (pos < B200_X_CAP) W.xsend[2 + pos] = ((unsigned long long)slot << 32) | (row | (code << B200_WV_ROW_BITS));
}
// one live row against the halfspace of one slot (rows a cut creates; the look-ahead kernel has a vectorised form)
B200_HD void wave_classify_row(const DevState &S, const WaveDev &W, u32 slot, u32 r, bool to_send = false)
{
	const CutParams &P = W.cur[slot];
	const int id = bit_test(S.ideal, r) ? 1 : 0;
	const double t = row_dot(S, P.h, r);
	double xinf = 0;
	for (int j = 0; j < S.d; j++) {
		const double a = wv_abs(S.coord[(size_t)j * S.cap_rows + r]);
		xinf = a > xinf ? a : xinf;
	}
	const u32 code = wave_code(t, id, P, xinf);
	if (code == 0xFFu) return;
	if (to_send) wave_send_append(W, slot, r, code);
	else wave_list_append(W, slot, r, code);
}

B200_HD void wave_reset_slot(CutCtl *c)
{
	u32 *w = (u32 *)c;
	for (u32 k = 0; k < sizeof(CutCtl) / 4; k++) w[k] = 0;
	c->min_strict_row = c->min_strict_slot = B200_NONE;
}

// halfspace of batch entry hs with the default callback's meaning (cone_polar, bslv_poly.c:30-39): vals.y >= -1, or >= 0
B200_HD void wave_make_params(const DevState &S, const WaveCtl *w, const double *vals, const unsigned char *ideal, u32 hs, CutParams &P)
{
	double hh = 0, h1 = 0;
	for (int j = 0; j < B200_MAXD; j++) {
		const double v = j < S.d ? vals[(size_t)hs * S.d + j] : 0.0;
		P.h[j] = v;
		hh = B200_ADD(hh, B200_MUL(v, v));
		h1 += wv_abs(v);
	}
	P.alpha = (ideal && ideal[hs]) ? 0.0 : -1.0;
	for (int id = 0; id < 2; id++) {
		const double thr = id ? 0.0 : P.alpha;
		P.hi[id] = B200_ADD(thr, 1e-9);
		P.mid[id] = B200_ADD(thr, 1.0e-2 * 1e-9);
		P.lo[id] = B200_SUB(thr, 1e-9);
	}
	P.hh = hh;
	P.h1 = h1;
	P.facet = w->facet0 + hs;
	P.batch_first = w->batch_first;
	P.seq = 0;
	P.zp_done = 0;
}

// ---------------------------------------------------------------- start of an iteration
// Decide whether a look-ahead pass runs and for which slots.  Works on a copy of the control block (shared
// memory on the device); the per-slot initialisation below is done by parallel threads afterwards.
B200_HD void wave_la_plan(WaveCtl &w, u32 nrows)
{
	w.n_la = 0;
	w.la_new = 0;
	if (w.halt) return;
	if (w.reclassify) {                  // every pending list is stale: rebuild them in this pass
		for (u32 p = 0; p < w.n_pending; p++) w.la[w.n_la++] = w.pending[p];
		w.reclassify = 0;
	}
	if (w.n_pending < w.refill_below || w.n_la) {
		u32 free_scan = 0;
		while (w.n_pending < B200_WAVE_SLOTS && w.next_hs < w.n_total) {
			while (free_scan < B200_WAVE_SLOTS && w.slot_hs[free_scan] != B200_NONE) free_scan++;
			if (free_scan >= B200_WAVE_SLOTS) break;
			const u32 slot = free_scan;
			w.slot_hs[slot] = w.next_hs++;
			w.pending[w.n_pending++] = slot;
			w.la_new |= 1u << w.n_la;
			w.la[w.n_la++] = slot;
		}
	}
	w.la_rows = nrows;
	if (w.n_la) {
		w.st_la_passes++;
		w.st_rows_scanned += nrows;
	}
}
// Is the pass just planned split across the ranks?  Every rank decides the same from the same numbers.
B200_HD void wave_shard_plan(WaveCtl &w, const WaveDev &W, u32 nrows)
{
	w.shard = 0;
	if (w.halt || !w.n_la) return;
	if (W.nranks > 1 && W.xsend && nrows >= W.shard_min_rows && !w.noshard_once) {
		w.shard = 1;
		w.xseq++;
		w.st_sharded++;
	}
	w.noshard_once = 0;
}
// this rank's share of the `ngroups` row groups of a sharded pass (ranges ascend with the rank)
B200_HD void wave_shard_range(const WaveDev &W, u32 ngroups, u32 &lo, u32 &hi)
{
	const u32 per = (ngroups + W.nranks - 1) / W.nranks;
	lo = per * W.rank < ngroups ? per * W.rank : ngroups;
	hi = lo + per < ngroups ? lo + per : ngroups;
}
// entry e of a received (or the own) record goes to the list of its slot
B200_HD void wave_merge_entry(const WaveDev &W, unsigned long long e)
{
	const u32 slot = (u32)(e >> 32), ent = (u32)e;
	const u32 pos = B200_ATOMIC_ADD(&W.ctl[slot].n_list, 1u);
	if (pos < B200_WAVE_LIST) W.list[(size_t)slot * B200_WAVE_LIST + pos] = ent;
}
// entry k of the pass: empty list; a halfspace that just received its slot also gets its parameters and its facet
B200_HD void wave_la_init(const DevState &S, const WaveDev &W, const WaveCtl &w, u32 k, const double *vals, const unsigned char *ideal)
{
	const u32 slot = w.la[k];
	wave_reset_slot(W.ctl + slot);
	if (!((w.la_new >> k) & 1u)) return;
	wave_make_params(S, &w, vals, ideal, w.slot_hs[slot], W.cur[slot]);
	S.facet_cnt[W.cur[slot].facet] = 0;
	S.facet_alive[W.cur[slot].facet] = 1;
}

// ---------------------------------------------------------------- wave formation
// Footprint marks: mark[row] = epoch << 5 | (31 - p) for the lowest pending position p whose cut may read or write
// the row, i.e. the rows of its list and their neighbours.  atomicMax keeps the current epoch and the lowest p.
B200_HD u32 wave_tag(u32 epoch, u32 p) { return (epoch << 5) | (31u - p); }

B200_HD void wave_mark_entry(const DevState &S, const WaveDev &W, u32 slot, u32 p, u32 epoch, u32 e)
{
	const u32 ent = W.list[(size_t)slot * B200_WAVE_LIST + e], row = ent & B200_WV_ROW_MASK, code = ent >> B200_WV_ROW_BITS;
	if (!bit_test(S.live, row)) return;                 // retired by an earlier cut
	if ((code & 3u) == CLS_ZP) B200_ATOMIC_OR(&W.wflag[p], 2u);   // may be projected in place (bslv_poly.c:666-674): runs alone
	const u32 tag = wave_tag(epoch, p);
	B200_ATOMIC_MAX(&W.mark[row], tag);
	const u32 off = S.adj_off[row], n = S.adj_len[row];
	for (u32 q = 0; q < n; q++) B200_ATOMIC_MAX(&W.mark[S.adj_pool[off + q]], tag);
}
B200_HD void wave_check_entry(const DevState &S, const WaveDev &W, u32 slot, u32 p, u32 epoch, u32 e)
{
	const u32 row = W.list[(size_t)slot * B200_WAVE_LIST + e] & B200_WV_ROW_MASK;
	if (!bit_test(S.live, row)) return;
	const u32 m = W.mark[row];
	if ((m >> 5) == epoch && (31u - (m & 31u)) < p) B200_ATOMIC_OR(&W.wflag[p], 1u);   // touches the footprint of an earlier pending cut
}
// candidates of this wave: the first `cand` pending slots
B200_HD u32 wave_candidates(const WaveCtl &w) { return w.n_pending < w.cand ? w.n_pending : w.cand; }

// flags[p]: bit0 = conflicts with an earlier candidate, bit1 = must run alone (ZERO+ row, list overflow)
B200_HD void wave_form_finish(WaveCtl &w, const u32 *flags)
{
	w.n_wave = 0;
	w.n_commit = 0;
	if (w.halt) return;
	if (w.n_pending == 0) {
		if (w.next_hs >= w.n_total) w.halt |= WH_DONE;
		return;
	}
	const u32 nc = wave_candidates(w);
	for (u32 p = 0; p < nc && w.n_wave < w.max_wave; p++) {
		const u32 fl = flags[p];
		if (fl & 2u) {                    // must run alone through the classic path; nothing overtakes it
			if (p == 0) { w.halt |= WH_SERIAL; w.halt_hs = w.slot_hs[w.pending[0]]; }
			break;
		}
		if (fl & 1u) {
			if (w.in_order) break;
			continue;                     // skipped: its footprint is marked, so no later cut that touches it joins
		}
		w.wave[w.n_wave++] = w.pending[p];
	}
	w.epoch++;
	w.st_waves++;
}

// ---------------------------------------------------------------- bases of the cuts of a wave
// After the sizes of every cut of the wave are known (new rows, incidence entries): where each cut appends, how
// many leading cuts of the wave can be carried out with the present capacities, and why the first one that
// cannot stops.  Every cluster derives the same plan from the same numbers.
struct WavePlan {
	u32 n_commit;
	u32 rows_base[B200_WAVE_MAXW], inc_base[B200_WAVE_MAXW], live_before[B200_WAVE_MAXW];
	u32 halt, halt_hs, need_rows, need_inc;
	u64 need_bits;
};
B200_HD void wave_gather_cut(const WaveDev &W, const WaveCtl &w, u32 q, WaveCut &o)
{
	const u32 slot = w.wave[q];
	const volatile CutCtl *c = W.ctl + slot;      // (volatile: other clusters of the running kernel may just have written it)
	o.status = c->status; o.n_new = c->n_new; o.inc_new = c->inc_new; o.n_minus = c->n_minus; o.n_zero = c->n_zero;
	o.n_pairs = c->n_pairs; o.n_surv = c->n_surv; o.adj_new = c->adj_new; o.live_before = c->n_live; o.padj_new = c->padj_new;
	o.facet = W.cur[slot].facet;
	o.hs = w.slot_hs[slot];
	o.slot = slot;
}
B200_HD void wave_plan(const WaveCtl &w, const WaveCut *cut, u32 nrows, u32 inc_used, u32 n_live, u32 cap_rows, u32 cap_inc, u64 cap_bits, WavePlan &pl)
{
	u32 rows = nrows, inc = inc_used, live = n_live;
	pl.n_commit = w.n_wave;
	pl.halt = 0;
	pl.halt_hs = B200_NONE;
	pl.need_rows = pl.need_inc = 0;
	pl.need_bits = 0;
	for (u32 q = 0; q < w.n_wave; q++) {
		const WaveCut &c = cut[q];
		pl.rows_base[q] = rows;
		pl.inc_base[q] = inc;
		pl.live_before[q] = live;
		if (c.status & ST_REDUNDANT) continue;
		// upper bound of the pair test's bit matrices: every facet a column
		const u32 mpad = (c.n_new + 63) & ~63u, wl_ub = (c.facet + 64) / 64;
		const u64 bits_ub = k4_words(wl_ub, mpad, wl_ub * 64);
		u32 bad = 0;
		if (c.status & (ST_NEED_BIG | ST_ERR_DEGENERATE | ST_OVF_PADJ)) bad = WH_SERIAL;
		else if ((u64)rows + c.n_new > cap_rows || (u64)rows + c.n_new > B200_WV_ROW_MASK || (u64)inc + c.inc_new > cap_inc || bits_ub > cap_bits) bad = WH_GROW;
		if (bad) {
			pl.n_commit = q;
			if (q == 0) {
				pl.halt = bad;
				pl.halt_hs = c.hs;
				pl.need_rows = rows + c.n_new;
				pl.need_inc = inc + c.inc_new;
				pl.need_bits = bits_ub;
			}
			break;
		}
		rows += c.n_new;
		inc += c.inc_new;
		live = live + c.n_new - (c.n_minus + c.n_zero);
	}
}

// ---------------------------------------------------------------- pair test results (wave path)
// The containment kernel does not compact the adjacent pairs into a second list (a position per pair means an
// atomic with a return value per block round): it flags the survivor in place; the adjacency build walks the
// survivor list.  The number of adjacent pairs is only needed as a sum (adjacency entries = PLUS neighbours + 2 * pairs).
#define B200_SURV_ADJ 0x80000000u
B200_HD void wave_flag_adjacent(const DevState &S, u32 s, u32 a, u32 b)
{
	S.surv_a[s] = a | B200_SURV_ADJ;
	B200_ATOMIC_ADD(&S.deg[a], 1u);
	B200_ATOMIC_ADD(&S.deg[b], 1u);
}
// sharded pair test: an adjacent pair found by this rank goes to its exchange record ...
B200_HD void wave_k4_send_pair(const WaveDev &W, u32 q, u32 a, u32 b)
{
	const u32 pos = B200_ATOMIC_ADD((u32 *)W.xksend, 1u);
	if (pos < B200_XK_CAP) W.xksend[4 + pos] = ((unsigned long long)q << 56) | ((unsigned long long)a << 28) | b;
}
// ... and every pair of every record is filed under its cut (pair list of the wave position, degrees, count)
B200_HD void wave_k4_merge_pair(const DevState &S0, const WaveDev &W, const WaveCtl &w, unsigned long long e)
{
	const u32 q = (u32)(e >> 56), a = (u32)(e >> 28) & 0x0FFFFFFFu, b = (u32)e & 0x0FFFFFFFu;
	const DevState S = wave_view(S0, W, w.wave[q], q);
	const u32 pos = B200_ATOMIC_ADD(&S.ctl->n_pairs, 1u);
	if (pos < S.cap_pairs) { S.pair_a[pos] = a; S.pair_b[pos] = b; }
	B200_ATOMIC_ADD(&S.deg[a], 1u);
	B200_ATOMIC_ADD(&S.deg[b], 1u);
}
B200_HD void wave_k4_record_reset(const WaveDev &W, u32 seq)
{	// header: count, exchange number, overflow flag, survivors needed
	W.xksend[0] = 0; W.xksend[1] = seq; W.xksend[2] = 0; W.xksend[3] = 0;
}
B200_HD void adj_pair_fill_surv(const DevState &S, u32 s)
{
	u32 a = S.surv_a[s];
	if (!(a & B200_SURV_ADJ)) return;
	a &= ~B200_SURV_ADJ;
	const u32 b = S.surv_b[s];
	const u32 nrows = S.ctl->nrows, used = S.ctl->adj_used;
	const u32 oa = used + S.adj_base[a] + S.new_padj_len[a], ob = used + S.adj_base[b] + S.new_padj_len[b];
	const u32 pa = B200_ATOMIC_ADD(&S.adj_fill[a], 1u), pb = B200_ATOMIC_ADD(&S.adj_fill[b], 1u);
	S.adj_pool[oa + pa] = nrows + b;
	S.adj_pool[ob + pb] = nrows + a;
}

// ---------------------------------------------------------------- commit (first kernel of the next iteration)
// On copies of the wave's and the polytope's control block; rc[q] receives the return code of wave position q.
B200_HD void wave_commit(WaveCtl &w, CutCtl &m, const WaveCut *cut, int dim, int *rc)
{
	const u64 d = (u64)dim;
	u64 evals = 0, bytes = 0, cuts = 0, red = 0, mn = 0, zr = 0, ed = 0, pt = 0, pr = 0;
	for (u32 q = 0; q < w.n_commit; q++) {
		const WaveCut &c = cut[q];
		if (c.status & ST_REDUNDANT) {
			const u64 N = c.live_before;          // (= the running live count: k_wave_tailB stores it for every carried-out position)
			evals += N;
			red++;
			bytes += N * (8 * d + 1);
			rc[q] = 1;
		} else {
			const u64 N = c.live_before;
			const u64 nm = c.n_minus, nz = c.n_zero, M = c.n_new, E = M - nz, Wd = ((u64)c.facet + 64) / 64, A = c.n_pairs;
			evals += N;
			cuts++;
			mn += nm;
			zr += nz;
			ed += E;
			pt += M * (M - (M ? 1 : 0)) / 2;
			pr += A;
			bytes += N * (8 * d + 1) + N + 4 * (nm + nz) + E * (24 * d + 24 * Wd) + nz * (16 * d + 16 * Wd) + 8 * M * Wd + 8 * A;
			m.n_live = m.n_live + c.n_new - (c.n_minus + c.n_zero);
			m.nrows += c.n_new;
			m.slot_cnt += c.n_new;
			m.inc_used += c.inc_new;
			m.adj_used += c.adj_new;
			rc[q] = 0;
		}
		w.slot_hs[w.wave[q]] = B200_NONE;
	}
	w.st_evals += evals; w.st_bytes += bytes; w.st_cuts += cuts; w.st_redundant += red; w.st_minus += mn; w.st_zero += zr;
	w.st_edge += ed; w.st_copies += zr; w.st_pair_tests += pt; w.st_pairs += pr;
	w.st_deferred += w.n_wave - w.n_commit;
	// the committed slots leave the pending list (order kept)
	u32 keep = 0;
	for (u32 p = 0; p < w.n_pending; p++) {
		const u32 slot = w.pending[p];
		if (w.slot_hs[slot] != B200_NONE) w.pending[keep++] = slot;
	}
	w.n_pending = keep;
	w.done_hs += w.n_commit;
	w.n_wave = w.n_commit = 0;
	if (w.done_hs >= w.n_total) w.halt |= WH_DONE;
	else if (m.nrows != m.n_live && m.nrows >= 4 * B200_TILE && m.nrows >= 2 * m.n_live) w.halt |= WH_COMPACT;
}
// adjacency entries cut q of the wave appends = PLUS neighbours of its new rows + two per adjacent pair; where each
// cut's block starts; whether the pool / the pair buffers hold it.  flags: 8 = pair buffers, 16 = adjacency pool.
B200_HD u32 wave_adj_plan(const WaveCut *cut, u32 n_commit, u32 adj_used, u32 cap_adj, u32 cap_pairs, u32 *adj_base, u32 *adj_new, u32 &need_adj, u32 &need_pairs)
{
	u32 base = adj_used, over = 0;
	for (u32 q = 0; q < n_commit; q++) {
		adj_base[q] = base;
		adj_new[q] = 0;
		if (cut[q].status & ST_REDUNDANT) continue;
		if (cut[q].n_surv > cap_pairs) over = over > cut[q].n_surv ? over : cut[q].n_surv;
		if (cut[q].n_pairs > cap_pairs) over = over > cut[q].n_pairs ? over : cut[q].n_pairs;   // (sharded pair test: the merged pair list)
		adj_new[q] = cut[q].padj_new + 2 * cut[q].n_pairs;
		base += adj_new[q];
	}
	need_adj = base;
	need_pairs = over;
	if (over) return 8u;
	if ((u64)base > cap_adj) return 16u;
	return 0;
}
B200_HD void wave_publish(const WaveDev &W, const WaveCtl &w, u32 nrows, u32 n_live)
{
	WaveProgress *p = W.progress;
	// (no fence between the words: the host uses `iter` only to pace its launches and re-reads everything from
	// device memory after a stream synchronise once it sees `halt`; a system-scope fence here would keep the last
	// kernel of every iteration alive for a PCIe round trip)
	p->done_hs = w.done_hs;
	p->nrows = nrows;
	p->n_live = n_live;
	p->iter = w.iter;
	p->halt = w.halt;
}
