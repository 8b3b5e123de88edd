// Per-work-item logic of the cut stages (K1..K5), shared between the sm_100a kernels
// (cut_kernels.cuh) and the host-side test double built with -DB200_EMULATE (tests only; the
// product library is never built with it and has no CPU path).
//
// Reference semantics follow SURVEY App. A; each function cites the bslv_poly.c lines it replaces.
// Floating point is the reference's operation order with separately rounded multiply and add
// (gcc -std=c99 => no contraction), so results are bit-identical to the CPU engine.
#pragma once
#include "cut_types.h"

#if defined(__CUDA_ARCH__)
#define B200_MUL(a, b) __dmul_rn((a), (b))
#define B200_ADD(a, b) __dadd_rn((a), (b))
#define B200_SUB(a, b) __dsub_rn((a), (b))
#define B200_DIV(a, b) __ddiv_rn((a), (b))
#define B200_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define B200_ATOMIC_SUB(p, v) atomicSub((p), (v))
#define B200_ATOMIC_OR(p, v) atomicOr((p), (v))
#define B200_ATOMIC_AND(p, v) atomicAnd((p), (v))
#define B200_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define B200_ATOMIC_EXCH(p, v) atomicExch((p), (v))
#define B200_ATOMIC_OR64(p, v) atomicOr((unsigned long long *)(p), (unsigned long long)(v))
#else // host compilation pass (must be built with -ffp-contract=off)
#define B200_MUL(a, b) ((a) * (b))
#define B200_ADD(a, b) ((a) + (b))
#define B200_SUB(a, b) ((a) - (b))
#define B200_DIV(a, b) ((a) / (b))
static inline u32 b200_fetch_add(u32 *p, u32 v) { u32 o = *p; *p = o + v; return o; }
static inline u32 b200_fetch_sub(u32 *p, u32 v) { u32 o = *p; *p = o - v; return o; }
static inline u32 b200_fetch_or(u32 *p, u32 v) { u32 o = *p; *p = o | v; return o; }
static inline u32 b200_fetch_and(u32 *p, u32 v) { u32 o = *p; *p = o & v; return o; }
static inline u32 b200_fetch_min(u32 *p, u32 v) { u32 o = *p; if (v < o) *p = v; return o; }
static inline u32 b200_exch(u32 *p, u32 v) { u32 o = *p; *p = v; return o; }
#define B200_ATOMIC_ADD(p, v) b200_fetch_add((p), (v))
#define B200_ATOMIC_SUB(p, v) b200_fetch_sub((p), (v))
#define B200_ATOMIC_OR(p, v) b200_fetch_or((p), (v))
#define B200_ATOMIC_AND(p, v) b200_fetch_and((p), (v))
#define B200_ATOMIC_MIN(p, v) b200_fetch_min((p), (v))
#define B200_ATOMIC_EXCH(p, v) b200_exch((p), (v))
#define B200_ATOMIC_OR64(p, v) (*(p) |= (v))
#endif

B200_HD bool bit_test(const u32 *w, u32 i) { return (w[i >> 5] >> (i & 31)) & 1u; }
B200_HD u32 popc64(u64 x)
{
#if defined(__CUDA_ARCH__)
	return (u32)__popcll(x);
#else
	return (u32)__builtin_popcountll(x);
#endif
}

// h . x for device row r, strict left-to-right from 0 (bslv_poly.c:123-125, 569-571, 593-595)
B200_HD double row_dot(const DevState &S, const double *h, u32 r)
{
	double s = B200_MUL(h[0], S.coord[r]);
	for (int j = 1; j < S.d; j++) s = B200_ADD(s, B200_MUL(h[j], S.coord[(size_t)j * S.cap_rows + r]));
	return s;
}

// A.2 classes from the three thresholds (bslv_poly.c:596, :666, :573)
B200_HD u8 class_of(double t, int ideal, const CutParams &P)
{
	if (t > P.hi[ideal]) return CLS_PLUS;
	if (t > P.mid[ideal]) return CLS_ZP;
	if (t > P.lo[ideal]) return CLS_ZERO;
	return CLS_MINUS;
}

// K1 for one row: class byte; bookkeeping of the trigger scan (bslv_poly.c:121-128)
B200_HD u8 classify_row(const DevState &S, const CutParams &P, u32 r, bool &strict, bool &zp)
{
	strict = zp = false;
	if (!bit_test(S.live, r)) return CLS_DEAD;
	int id = bit_test(S.ideal, r) ? 1 : 0;
	double t = row_dot(S, P.h, r);
	strict = t < P.lo[id];
	u8 c = class_of(t, id, P);
	zp = (c == CLS_ZP);
	return c;
}

B200_HD bool is_visited_class(u8 c) { return c == CLS_ZERO || c == CLS_MINUS; }

// ZERO+ closure step for visited-list entry i (bslv_poly.c:666-674): a vertex with
// thr+1e-11 < h.x <= thr+1e-9 that neighbours a visited vertex is projected onto the hyperplane in
// place and then treated like any other visited vertex.  Returns true if it activated the entry.
B200_HD bool zp_activate(const DevState &S, const CutParams &P, u32 i)
{
	u32 v = S.vis[i];
	if (S.cls[v] != CLS_ZP) return false;
	bool touch = false;
	for (u32 q = 0, off = S.adj_off[v], n = S.adj_len[v]; q < n && !touch; q++)
		touch = is_visited_class(S.cls[S.adj_pool[off + q]]);
	if (!touch) return false;
	int id = bit_test(S.ideal, v) ? 1 : 0;
	double thr = id ? 0.0 : P.alpha;
	if (!P.zp_done) {                    // (a rerun after a bail-out finds the row already projected by the first attempt)
		double mu = B200_DIV(B200_SUB(row_dot(S, P.h, v), thr), P.hh);
		for (int j = 0; j < S.d; j++) {
			double *x = S.coord + (size_t)j * S.cap_rows + v;
			*x = B200_SUB(*x, B200_MUL(mu, P.h[j]));
		}
	}
	double t = row_dot(S, P.h, v);
	S.cls[v] = (t > P.lo[id]) ? CLS_ZERO : CLS_MINUS;   // re-test of bslv_poly.c:573 after :674
	B200_ATOMIC_ADD(&S.ctl->n_zp_projected, 1u);
	return true;
}

// Short lists (simple vertices have exactly d entries) are fetched with independent loads and
// compared in registers: a sorted merge over global memory is a chain of dependent L2 round trips.
#define B200_SHORT 8
B200_HD void load_short(const u32 *p, u32 n, u32 r[B200_SHORT])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 t = 0; t < B200_SHORT; t++) r[t] = t < n ? p[t] : B200_NONE;
}
// bit t of the result <=> a[t] also occurs in b   (both lists at most B200_SHORT long, entries distinct)
B200_HD u32 common_mask_short(const u32 ra[B200_SHORT], u32 na, const u32 rb[B200_SHORT])
{
	u32 m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 t = 0; t < B200_SHORT; t++) {
		bool hit = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 u = 0; u < B200_SHORT; u++) hit |= (rb[u] == ra[t]);
		if (t < na && hit) m |= 1u << t;
	}
	return m;
}
B200_HD u32 popc32(u32 x)
{
#if defined(__CUDA_ARCH__)
	return (u32)__popc(x);
#else
	return (u32)__builtin_popcount(x);
#endif
}

// |A n B| of two sorted lists
B200_HD u32 isect_count(const u32 *a, u32 na, const u32 *b, u32 nb)
{
	if (na <= B200_SHORT && nb <= B200_SHORT) {
		u32 ra[B200_SHORT], rb[B200_SHORT];
		load_short(a, na, ra);
		load_short(b, nb, rb);
		return popc32(common_mask_short(ra, na, rb));
	}
	u32 i = 0, j = 0, n = 0;
	while (i < na && j < nb) {
		u32 x = a[i], y = b[j];
		n += (x == y);
		i += (x <= y);
		j += (y <= x);
	}
	return n;
}

// On-plane vertex whose incidence list is longer than the shared-facet mask (B200_MAXINC positions): does facet x of
// inc(v) also lie on one of v's PLUS neighbours (bslv_poly.c:637-652)?  The mask-free form: a binary search in each
// PLUS neighbour's sorted list.  Rare and slow by design -- it is what keeps a recession direction of an upper image
// (on thousands of facets after thousands of Benson cuts) from being an error.
B200_HD bool zero_keeps_facet(const DevState &S, u32 v, u32 x)
{
	for (u32 q = 0, off = S.adj_off[v], n = S.adj_len[v]; q < n; q++) {
		const u32 k = S.adj_pool[off + q];
		if (S.cls[k] != CLS_PLUS) continue;
		const u32 *ik = S.inc_pool + S.inc_off[k];
		u32 lo = 0, hi = S.inc_len[k];
		while (lo < hi) {
			const u32 mid = (lo + hi) >> 1;
			if (ik[mid] < x) lo = mid + 1;
			else hi = mid;
		}
		if (lo < S.inc_len[k] && ik[lo] == x) return true;
	}
	return false;
}
B200_HD u32 zero_long_count(const DevState &S, u32 v)
{
	const u32 *iv = S.inc_pool + S.inc_off[v];
	u32 n = 0;
	for (u32 a = 0, niv = S.inc_len[v]; a < niv; a++) n += zero_keeps_facet(S, v, iv[a]) ? 1u : 0u;
	return n;
}

// The fast form of the same: visited entry i owns row i of a facet bitmap (DevState::zlong) that the PLUS neighbours'
// lists are OR-ed into -- linear in the list lengths and parallel over the neighbours.  Available when the bitmap
// covers every facet id and the entry has a row; the copy stage clears the row again.
B200_HD bool zlong_usable(const DevState &S, u32 i) { return S.zlong && i < B200_VIS_MAX && S.cur->facet < S.zlong_words * 64u; }
B200_HD void zlong_add_list(const DevState &S, u32 i, u32 k)
{
	u64 *row = S.zlong + (size_t)i * S.zlong_words;
	const u32 *ik = S.inc_pool + S.inc_off[k];
	for (u32 b = 0, nik = S.inc_len[k]; b < nik; b++) B200_ATOMIC_OR64(&row[ik[b] >> 6], (u64)1 << (ik[b] & 63));
}
B200_HD bool zlong_test(const DevState &S, u32 i, u32 x) { return (S.zlong[(size_t)i * S.zlong_words + (x >> 6)] >> (x & 63)) & 1; }
B200_HD u32 zlong_count(const DevState &S, u32 i, u32 v)
{
	const u32 *iv = S.inc_pool + S.inc_off[v];
	u32 n = 0;
	for (u32 a = 0, niv = S.inc_len[v]; a < niv; a++) n += zlong_test(S, i, iv[a]) ? 1u : 0u;
	return n;
}

// K3a: how many new rows / incidence entries / PLUS neighbours visited entry i produces
// (sizes of what bslv_poly.c:573-588 and :597-665 append)
B200_HD void count_outputs(const DevState &S, u32 i)
{
	u32 v = S.vis[i];
	u8 c = S.cls[v];
	u32 n_out = 0, inc_sz = 0, nplus = 0;
	if (is_visited_class(c)) {
		const u32 *iv = S.inc_pool + S.inc_off[v];
		const u32 niv = S.inc_len[v];
		u64 mask[B200_MAXINC / 64] = {0};
		const bool zlong = c == CLS_ZERO && niv > B200_MAXINC;       // longer than the mask: counted without it below
		for (u32 q = 0, off = S.adj_off[v], n = S.adj_len[v]; q < n; q++) {
			u32 k = S.adj_pool[off + q];
			if (S.cls[k] != CLS_PLUS) continue;
			nplus++;
			const u32 *ik = S.inc_pool + S.inc_off[k];
			const u32 nik = S.inc_len[k];
			if (c == CLS_MINUS)
				inc_sz += 1 + isect_count(iv, niv, ik, nik);
			else if (zlong) {
				if (zlong_usable(S, i)) zlong_add_list(S, i, k);
			} else {
				u32 a = 0, b = 0;
				while (a < niv && a < B200_MAXINC && b < nik) {
					u32 x = iv[a], y = ik[b];
					if (x == y) mask[a >> 6] |= (u64)1 << (a & 63);
					a += (x <= y);
					b += (y <= x);
				}
			}
		}
		if (c == CLS_ZERO) {
			n_out = 1;
			inc_sz = 1;
			if (zlong) inc_sz += zlong_usable(S, i) ? zlong_count(S, i, v) : zero_long_count(S, v);
			else
				for (int w = 0; w < B200_MAXINC / 64; w++) {
					u64 m = mask[w];
					while (m) { m &= m - 1; inc_sz++; }
				}
		} else
			n_out = nplus;
	}
	S.cnt3[3 * (size_t)i + 0] = n_out;
	S.cnt3[3 * (size_t)i + 1] = inc_sz;
	S.cnt3[3 * (size_t)i + 2] = nplus;
}

B200_HD void set_bit_atomic(u32 *w, u32 i) { B200_ATOMIC_OR(&w[i >> 5], 1u << (i & 31)); }
B200_HD void clr_bit_atomic(u32 *w, u32 i) { B200_ATOMIC_AND(&w[i >> 5], ~(1u << (i & 31))); }

// neighbour k of the dying row v now neighbours `nw` instead (bslv_poly.c:628-632)
B200_HD void rewire(const DevState &S, u32 k, u32 v, u32 nw)
{
	u32 off = S.adj_off[k], n = S.adj_len[k];
	if (n <= B200_SHORT) {
		u32 r[B200_SHORT];
		load_short(S.adj_pool + off, n, r);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 q = 0; q < B200_SHORT; q++)
			if (q < n && r[q] == v) { S.adj_pool[off + q] = nw; return; }
		return;
	}
	for (u32 q = 0; q < n; q++)
		if (S.adj_pool[off + q] == v) { S.adj_pool[off + q] = nw; return; }
}

// K4 column relabelling (see k4_assign_columns below): the first toucher of a facet in a cut allocates the
// facet's column of the cut's bit matrix; epoch = new facet id + 1 tags the cut.
B200_HD void k4_assign_one(const DevState &S, u32 fc, u32 epoch)
{
	if (B200_ATOMIC_EXCH(&S.facet_epoch[fc], epoch) != epoch) S.facet_local[fc] = B200_ATOMIC_ADD(&S.ctl->n_local, 1u);
}
// same for the masked entries of a short list: the exchanges are independent and go out together, the fresh
// columns are taken with one counter update
B200_HD void k4_assign_short(const DevState &S, const u32 fc[B200_SHORT], u32 mask, u32 epoch)
{
	u32 old[B200_SHORT];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 t = 0; t < B200_SHORT; t++) old[t] = ((mask >> t) & 1u) ? B200_ATOMIC_EXCH(&S.facet_epoch[fc[t]], epoch) : epoch;
	u32 fresh = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 t = 0; t < B200_SHORT; t++) fresh |= (old[t] != epoch ? 1u : 0u) << t;
	if (!fresh) return;
	u32 col = B200_ATOMIC_ADD(&S.ctl->n_local, popc32(fresh));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 t = 0; t < B200_SHORT; t++)
		if ((fresh >> t) & 1u) S.facet_local[fc[t]] = col++;
}

// ---- K3b pieces, shared by the vertex-serial form (emit_outputs) and the half-edge-parallel form

// new vertex on the edge (v MINUS, k PLUS), SURVEY A.3 / bslv_poly.c:597-627 + incidence :634-665.
// j = index among this cut's new rows, ipos = where its incidence list goes, pslot = its padj slot.
// The two halves below are independent of each other (geometry + adjacency / incidence + facet counts), so the
// tail kernels give them to different warps; the bodies are chains of dependent memory round trips if written
// naively, so in each the loads are grouped in waves and issued before the first store (stores end the compiler's
// freedom to hoist loads).
B200_HD void emit_edge_geom(const DevState &S, const CutParams &P, u32 v, u32 k, u32 j, u32 pslot)
{
	const CutCtl *ctl = S.ctl;
	const size_t cap = S.cap_rows;
	const int d = S.d;
	// ---- wave 1: everything addressed by v and k alone
	const u32 nw = ctl->nrows + j, slot = ctl->slot_cnt + j;
	const bool v_ideal = bit_test(S.ideal, v), k_ideal = bit_test(S.ideal, k);
	const u32 aoff = S.adj_off[k], an = S.adj_len[k];
	const bool both = k_ideal && v_ideal, none = !k_ideal && !v_ideal;
	double base[B200_MAXD], dir[B200_MAXD];
	for (int t0 = 0; t0 < d; t0 += 8) {            // 16 independent coordinate loads per round
		double cv[8], ck[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int u = 0; u < 8; u++) {
			cv[u] = t0 + u < d ? S.coord[(t0 + u) * cap + v] : 0.0;
			ck[u] = t0 + u < d ? S.coord[(t0 + u) * cap + k] : 0.0;
		}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (int u = 0; u < 8; u++)
			if (t0 + u < d) {
				base[t0 + u] = k_ideal ? cv[u] : ck[u];          // base point
				double dv = k_ideal ? ck[u] : cv[u];              // direction source
				if (both) dv = B200_SUB(dv, cv[u]);
				else if (none) dv = B200_SUB(dv, ck[u]);
				dir[t0 + u] = dv;
			}
	}
	// ---- wave 2: k's neighbours
	const bool short_adj = an <= B200_SHORT;
	u32 ra[B200_SHORT];
	load_short(S.adj_pool + aoff, short_adj ? an : 0, ra);
	// ---- arithmetic: the reference's operation order (bslv_poly.c:597-627)
	double hb = B200_MUL(P.h[0], base[0]), hd = B200_MUL(P.h[0], dir[0]);
	for (int t = 1; t < d; t++) {
		hb = B200_ADD(hb, B200_MUL(P.h[t], base[t]));
		hd = B200_ADD(hd, B200_MUL(P.h[t], dir[t]));
	}
	const double mu = B200_DIV(B200_SUB(both ? 0.0 : P.alpha, hb), hd);
	// ---- stores
	for (int t = 0; t < d; t++) S.coord[t * cap + nw] = B200_ADD(base[t], B200_MUL(mu, dir[t]));
	if (both) set_bit_atomic(S.ideal, nw);
	set_bit_atomic(S.live, nw);
	S.cls[nw] = CLS_PLUS;                     // invariant: every live row reads PLUS between cuts
	S.row_slot[nw] = slot;
	S.new_parent[j] = B200_NONE;
	S.root[nw] = B200_NONE;
	S.deg[j] = 0;
	S.adj_fill[j] = 0;
	S.new_padj_off[j] = pslot;
	S.new_padj_len[j] = 1;
	S.padj[pslot] = k;
	// neighbour k of the dying row v now neighbours nw (bslv_poly.c:628-632)
	if (short_adj) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 q = 0; q < B200_SHORT; q++)
			if (q < an && ra[q] == v) { S.adj_pool[aoff + q] = nw; break; }
	} else
		rewire(S, k, v, nw);
}
// incidence {f} u (inc(k) n inc(v)), sorted; f is the largest facet id so far.  Each facet the new row lies on
// also gets its column of this cut's K4 bit matrix here (first toucher allocates it).
B200_HD void emit_edge_inc(const DevState &S, const CutParams &P, u32 v, u32 k, u32 j, u32 ipos)
{
	const u32 f = P.facet, nw = S.ctl->nrows + j;
	const u32 iov = S.inc_off[v], iok = S.inc_off[k], niv = S.inc_len[v], nik = S.inc_len[k];
	const bool short_inc = niv <= B200_SHORT && nik <= B200_SHORT;
	u32 rv[B200_SHORT], rk[B200_SHORT];
	load_short(S.inc_pool + iov, short_inc ? niv : 0, rv);
	load_short(S.inc_pool + iok, short_inc ? nik : 0, rk);
	u32 w = ipos;
	// f goes where it keeps the list sorted: last when cuts run in halfspace order, possibly earlier when a wave
	// carried out a later halfspace first
	bool placed = false;
	if (short_inc) {
		const u32 m = common_mask_short(rv, niv, rk);
		k4_assign_short(S, rv, m, f + 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 t = 0; t < B200_SHORT; t++)
			if ((m >> t) & 1u) {
				if (!placed && rv[t] > f) { S.inc_pool[w++] = f; placed = true; }
				S.inc_pool[w++] = rv[t];
				B200_ATOMIC_ADD(&S.facet_cnt[rv[t]], 1u);
			}
	} else {
		const u32 *iv = S.inc_pool + iov, *ik = S.inc_pool + iok;
		u32 a = 0, b = 0;
		while (a < niv && b < nik) {
			const u32 x = iv[a], y = ik[b];
			if (x == y) {
				if (!placed && x > f) { S.inc_pool[w++] = f; placed = true; }
				S.inc_pool[w++] = x;
				B200_ATOMIC_ADD(&S.facet_cnt[x], 1u);
				k4_assign_one(S, x, f + 1);
			}
			a += (x <= y);
			b += (y <= x);
		}
	}
	if (!placed) S.inc_pool[w++] = f;         // facet_cnt[f] is set once to n_new by the plan stage
	S.inc_off[nw] = ipos;
	S.inc_len[nw] = w - ipos;
}
B200_HD void emit_edge_vertex(const DevState &S, const CutParams &P, u32 v, u32 k, u32 j, u32 ipos, u32 pslot)
{
	emit_edge_geom(S, P, v, k, j, pslot);
	emit_edge_inc(S, P, v, k, j, ipos);
}

// which elements of inc(v) also lie in inc(k): bit a of mask <=> iv[a] in inc(k)   (bslv_poly.c:634-652)
B200_HD void shared_facet_mask(const DevState &S, u32 v, u32 k, u64 mask[B200_MAXINC / 64])
{
	const u32 *iv = S.inc_pool + S.inc_off[v], *ik = S.inc_pool + S.inc_off[k];
	const u32 niv = S.inc_len[v], nik = S.inc_len[k];
	if (niv <= B200_SHORT && nik <= B200_SHORT) {
		u32 rv[B200_SHORT], rk[B200_SHORT];
		load_short(iv, niv, rv);
		load_short(ik, nik, rk);
		mask[0] |= (u64)common_mask_short(rv, niv, rk);
		return;
	}
	u32 a = 0, b = 0;
	while (a < niv && a < B200_MAXINC && b < nik) {
		const u32 x = iv[a], y = ik[b];
		if (x == y) mask[a >> 6] |= (u64)1 << (a & 63);
		a += (x <= y);
		b += (y <= x);
	}
}

// the copy of an on-plane vertex v (bslv_poly.c:573-588): same coordinates, incidence = the masked
// part of inc(v) plus the new facet; its PLUS neighbours were attached separately.
B200_HD void emit_copy_row(const DevState &S, const CutParams &P, u32 i, u32 v, u32 j, u32 ipos, u32 pslot, u32 nplus,
                           const u64 mask[B200_MAXINC / 64])
{
	const CutCtl *ctl = S.ctl;
	const size_t cap = S.cap_rows;
	const u32 nw = ctl->nrows + j, f = P.facet;
	for (int t = 0; t < S.d; t++) S.coord[t * cap + nw] = S.coord[t * cap + v];
	if (bit_test(S.ideal, v)) set_bit_atomic(S.ideal, nw);
	set_bit_atomic(S.live, nw);
	S.cls[nw] = CLS_PLUS;
	S.row_slot[nw] = ctl->slot_cnt + j;
	S.new_parent[j] = S.row_slot[v];
	S.root[nw] = S.row_slot[v] < P.batch_first ? S.row_slot[v] : S.root[v];
	S.deg[j] = 0;
	S.adj_fill[j] = 0;
	S.new_padj_off[j] = pslot;
	S.new_padj_len[j] = nplus;
	const u32 *iv = S.inc_pool + S.inc_off[v];
	const u32 niv = S.inc_len[v];
	u32 w = ipos;
	bool placed = false;                      // f keeps the list sorted (see emit_edge_inc)
	const bool zlong = niv > B200_MAXINC;     // (then the mask was not built: the facet bitmap of entry i, or a direct test)
	const bool zbits = zlong && zlong_usable(S, i);
	for (u32 a = 0; a < niv; a++)
		if (zbits ? zlong_test(S, i, iv[a]) : zlong ? zero_keeps_facet(S, v, iv[a]) : (bool)((mask[a >> 6] >> (a & 63)) & 1)) {
			if (!placed && iv[a] > f) { S.inc_pool[w++] = f; placed = true; }
			S.inc_pool[w++] = iv[a];
			B200_ATOMIC_ADD(&S.facet_cnt[iv[a]], 1u);
			k4_assign_one(S, iv[a], f + 1);
		}
	if (!placed) S.inc_pool[w++] = f;         // facet_cnt[f] is set once to n_new by the plan stage
	S.inc_off[nw] = ipos;
	S.inc_len[nw] = w - ipos;
	if (zbits) {                              // the bitmap row reads zero again between cuts
		u64 *row = S.zlong + (size_t)i * S.zlong_words;
		for (u32 x = 0; x < S.zlong_words; x++) row[x] = 0;
	}
}

// retire visited row v (bslv_poly.c:568, 679-688, 697-705): its facets lose one vertex
B200_HD void retire_row(const DevState &S, u32 v, u32 i)
{
	const u32 off = S.inc_off[v], n = S.inc_len[v], slot = S.row_slot[v];
	if (n <= B200_SHORT) {
		u32 r[B200_SHORT];
		load_short(S.inc_pool + off, n, r);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 a = 0; a < B200_SHORT; a++)
			if (a < n) B200_ATOMIC_SUB(&S.facet_cnt[r[a]], 1u);
	} else {
		const u32 *iv = S.inc_pool + off;
		for (u32 a = 0; a < n; a++) B200_ATOMIC_SUB(&S.facet_cnt[iv[a]], 1u);
	}
	clr_bit_atomic(S.live, v);
	S.dead_slots[i] = slot;
}

// K3b + K5 for visited entry i, vertex-serial form (multi-kernel path and host test double)
B200_HD void emit_outputs(const DevState &S, const CutParams &P, u32 i)
{
	const u32 v = S.vis[i];
	const u8 c = S.cls[v];
	if (!is_visited_class(c)) { S.dead_slots[i] = B200_NONE; return; }
	CutCtl *ctl = S.ctl;
	const u32 jrow = S.base3[3 * (size_t)i + 0];
	u32 ipos = ctl->inc_used + S.base3[3 * (size_t)i + 1];
	const u32 ppos = S.base3[3 * (size_t)i + 2];
	const u32 aoff = S.adj_off[v], an = S.adj_len[v];
	u32 np = 0;
	if (c == CLS_ZERO) {
		const u32 nw = ctl->nrows + jrow;
		u64 mask[B200_MAXINC / 64] = {0};
		for (u32 q = 0; q < an; q++) {
			const u32 k = S.adj_pool[aoff + q];
			if (S.cls[k] != CLS_PLUS) continue;
			S.padj[ppos + np++] = k;
			rewire(S, k, v, nw);
			if (S.inc_len[v] <= B200_MAXINC) shared_facet_mask(S, v, k, mask);
		}
		emit_copy_row(S, P, i, v, jrow, ipos, ppos, np, mask);
		B200_ATOMIC_ADD(&ctl->n_zero, 1u);
	} else {
		for (u32 q = 0; q < an; q++) {
			const u32 k = S.adj_pool[aoff + q];
			if (S.cls[k] != CLS_PLUS) continue;
			emit_edge_vertex(S, P, v, k, jrow + np, ipos, ppos + np);
			ipos += S.inc_len[ctl->nrows + jrow + np];
			np++;
		}
		B200_ATOMIC_ADD(&ctl->n_minus, 1u);
	}
	retire_row(S, v, i);
}

// ---- half-edge-parallel form (tail kernels): one work item per (visited vertex, adjacency slot)
B200_HD void he_owner_fill(const DevState &S, u32 i)
{
	for (u32 e = S.he_off[i]; e < S.he_off[i + 1]; e++) S.he_own[e] = i;
}
// half-edge e of visited entry i (row v, first half-edge off_i); the caller knows the owner
B200_HD void he_eval_at(const DevState &S, u32 e, u32 i, u32 v, u32 off_i)
{
	// loads addressed by v alone go out together with the adjacency lookup, those addressed by k follow in one wave
	const u32 aoff = S.adj_off[v];
	const u8 cv = S.cls[v];
	const u32 iov = S.inc_off[v], niv = S.inc_len[v];
	const u32 k = S.adj_pool[aoff + (e - off_i)];
	const u8 ck = S.cls[k];
	const u32 iok = S.inc_off[k], nik = S.inc_len[k];
	const bool plus = ck == CLS_PLUS;
	u32 inc = 0;
	if (plus) {
		if (cv == CLS_MINUS) {
			inc = 1 + isect_count(S.inc_pool + iov, niv, S.inc_pool + iok, nik);
		} else if (niv > B200_MAXINC) {
			if (zlong_usable(S, i)) zlong_add_list(S, i, k);
		} else {
			u64 mask[B200_MAXINC / 64] = {0};
			shared_facet_mask(S, v, k, mask);
			const int nw = (int)((niv + 63) / 64) < B200_MAXINC / 64 ? (int)((niv + 63) / 64) : B200_MAXINC / 64;
			for (int w = 0; w < nw; w++)
				if (mask[w]) B200_ATOMIC_OR64(&S.zmask[(size_t)i * (B200_MAXINC / 64) + w], mask[w]);
		}
	}
	S.he_k[e] = k;
	S.he_flag[e] = plus ? 1 : 0;
	S.he_inc[e] = inc;
}
// sizes of what visited entry i (row v, class c, half-edges [e0, e1)) produces: out[0] new rows, out[1]
// incidence entries, out[2] PLUS neighbours; also each half-edge's position among the PLUS ones.
// (Always true now: on-plane vertices with lists longer than the mask are counted without it.)
B200_HD bool he_count_core(const DevState &S, u32 i, u32 v, u8 c, u32 e0, u32 e1, u32 out[3])
{
	u32 n_out = 0, inc_sz = 0, nplus = 0;
	bool ok = true;
	// a row that is not cut has no half-edges (e0 == e1), so the loop needs no class test and its loads do not
	// wait for the class
	if (e1 - e0 <= 12) {                                        // usual degree: all loads in flight before the first store
		u32 fl[12], ic[12];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 t = 0; t < 12; t++) {
			fl[t] = e0 + t < e1 ? S.he_flag[e0 + t] : 0;
			ic[t] = e0 + t < e1 ? S.he_inc[e0 + t] : 0;
		}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 t = 0; t < 12; t++)
			if (e0 + t < e1) {
				S.he_rank[e0 + t] = nplus;
				S.he_incpre[e0 + t] = inc_sz;
				nplus += fl[t];
				inc_sz += ic[t];
			}
	} else
		for (u32 e = e0; e < e1; e++) {
			S.he_rank[e] = nplus;
			S.he_incpre[e] = inc_sz;
			nplus += S.he_flag[e];
			inc_sz += S.he_inc[e];
		}
	if (is_visited_class(c)) {
		if (c == CLS_ZERO) {
			n_out = 1;
			inc_sz = 1;
			if (S.inc_len[v] > B200_MAXINC) inc_sz += zlong_usable(S, i) ? zlong_count(S, i, v) : zero_long_count(S, v);
			else {
				const int nw = (int)((S.inc_len[v] + 63) / 64);
				for (int w = 0; w < nw; w++) inc_sz += popc64(S.zmask[(size_t)i * (B200_MAXINC / 64) + w]);
			}
		} else
			n_out = nplus;
	} else
		inc_sz = nplus = 0;
	out[0] = n_out;
	out[1] = inc_sz;
	out[2] = nplus;
	return ok;
}
B200_HD void he_eval(const DevState &S, u32 e)
{
	const u32 i = S.he_own[e];
	he_eval_at(S, e, i, S.vis[i], S.he_off[i]);
}
B200_HD void he_count(const DevState &S, u32 i)
{
	const u32 v = S.vis[i];
	u32 out[3];
	if (!he_count_core(S, i, v, S.cls[v], S.he_off[i], S.he_off[i + 1], out)) B200_ATOMIC_OR(&S.ctl->status, (u32)ST_ERR_DEGENERATE);
	S.cnt3[3 * (size_t)i + 0] = out[0];
	S.cnt3[3 * (size_t)i + 1] = out[1];
	S.cnt3[3 * (size_t)i + 2] = out[2];
}
// part 0: geometry, bookkeeping and rewiring of the new row; part 1: its incidence list; 2: both
B200_HD void he_emit_part(const DevState &S, const CutParams &P, u32 e, int part)
{
	const u32 flag = S.he_flag[e], i = S.he_own[e], k = S.he_k[e], rank = S.he_rank[e], incpre = S.he_incpre[e];
	if (!flag) return;
	const u32 v = S.vis[i];
	const u32 jrow = S.base3[3 * (size_t)i + 0], ibase = S.ctl->inc_used + S.base3[3 * (size_t)i + 1], ppos = S.base3[3 * (size_t)i + 2];
	if (S.cls[v] == CLS_MINUS) {
		if (part != 1) emit_edge_geom(S, P, v, k, jrow + rank, ppos + rank);
		if (part != 0) emit_edge_inc(S, P, v, k, jrow + rank, ibase + incpre);
	} else if (part != 1) {
		S.padj[ppos + rank] = k;
		rewire(S, k, v, S.ctl->nrows + jrow);
	}
}
B200_HD void he_emit(const DevState &S, const CutParams &P, u32 e) { he_emit_part(S, P, e, 2); }
B200_HD void he_finish_vertex(const DevState &S, const CutParams &P, u32 i)
{
	const u32 v = S.vis[i];
	const u8 c = S.cls[v];
	if (!is_visited_class(c)) { S.dead_slots[i] = B200_NONE; return; }
	if (c == CLS_ZERO) {
		emit_copy_row(S, P, i, v, S.base3[3 * (size_t)i + 0], S.ctl->inc_used + S.base3[3 * (size_t)i + 1], S.base3[3 * (size_t)i + 2],
		              S.cnt3[3 * (size_t)i + 2], &S.zmask[(size_t)i * (B200_MAXINC / 64)]);
	}
	retire_row(S, v, i);                      // n_minus / n_zero were counted by the plan stage
}

// redundant halfspace: nothing is cut, the non-PLUS rows K1 marked go back to PLUS
B200_HD void reset_class(const DevState &S, u32 i) { S.cls[S.vis[i]] = CLS_PLUS; }

// after all counts are final: a facet without live vertices dies (clean rule; the reference's
// variant, bslv_poly.c:686-687/:705, leaves order-dependent ghosts -- SURVEY section 0)
B200_HD void collect_dead_facets(const DevState &S, u32 i)
{
	const u32 v = S.vis[i];
	const u32 ds = S.dead_slots[i], off = S.inc_off[v], n = S.inc_len[v];
	if (ds == B200_NONE) { S.cls[v] = CLS_PLUS; return; }   // a ZERO+ row nobody reached stays as it is
	if (n <= B200_SHORT) {
		u32 r[B200_SHORT], cnt[B200_SHORT];
		load_short(S.inc_pool + off, n, r);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 a = 0; a < B200_SHORT; a++) cnt[a] = a < n ? S.facet_cnt[r[a]] : 1u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 a = 0; a < B200_SHORT; a++)
			if (cnt[a] == 0 && B200_ATOMIC_EXCH(&S.facet_alive[r[a]], 0u) == 1u)
				S.dead_facets[B200_ATOMIC_ADD(&S.ctl->n_dead_facets, 1u)] = r[a];
		return;
	}
	const u32 *iv = S.inc_pool + off;
	for (u32 a = 0; a < n; a++) {
		u32 fc = iv[a];
		if (S.facet_cnt[fc] == 0 && B200_ATOMIC_EXCH(&S.facet_alive[fc], 0u) == 1u)
			S.dead_facets[B200_ATOMIC_ADD(&S.ctl->n_dead_facets, 1u)] = fc;
	}
}

// ---- K4, bitset form (north_star (3)): the incidence lists of the M new rows are re-coded as rows of a
// bit matrix over the facets they actually touch (the new facet itself is common to all and left
// out, so the reference's |mutual| >= d-1, bslv_poly.c:484, becomes popcount >= d-2).

// column relabelling: first toucher of a facet in this cut allocates its column
B200_HD void k4_assign_columns(const DevState &S, u32 j)
{
	// (emit_edge_vertex / emit_copy_row have done this for the rows they created: every exchange below then
	// finds the tag in place; the stage remains for paths that build rows differently)
	const u32 r = S.ctl->nrows + j, f = S.cur->facet, epoch = f + 1;
	const u32 off = S.inc_off[r], n = S.inc_len[r];
	if (n <= B200_SHORT) {
		u32 l[B200_SHORT];
		load_short(S.inc_pool + off, n, l);
		u32 m = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 q = 0; q < B200_SHORT; q++) m |= (q < n && l[q] != f ? 1u : 0u) << q;
		k4_assign_short(S, l, m, epoch);
		return;
	}
	const u32 *l = S.inc_pool + off;
	for (u32 q = 0; q < n; q++) {
		const u32 fc = l[q];
		if (fc == f) continue;
		k4_assign_one(S, fc, epoch);
	}
}
// K4 keeps two packed forms of the incidence of the M new rows over the L facets they touch:
//   rows : bits[w*mpad + x]            bit c of word w  <=> row x lies on local facet 64w+c   (filter: AND + POPC of two rows)
//   cols : tbits[col*(mpad/64) + x/64] bit x%64         <=> row x lies on local facet col     (containment: AND of columns)
// tbits starts right after the row matrix in the same buffer.
B200_HD u64 k4_words(u32 wl, u32 mpad, u32 n_local) { return (u64)wl * mpad + (u64)n_local * (mpad / 64); }
B200_HD u64 *k4_tbits(const DevState &S, u32 wl, u32 mpad) { return S.bits + (size_t)wl * mpad; }

B200_HD void k4_plan(const DevState &S)
{
	CutCtl *c = S.ctl;
	c->wl = (c->n_local + 63) / 64;
	c->mpad = (c->n_new + 63) & ~63u;
	if (k4_words(c->wl, c->mpad, c->n_local) > S.cap_bits) c->status |= ST_OVF_BITS;
}
B200_HD void k4_build_row_at(const DevState &S, u32 j, u32 nrows, u32 f, u32 wl, u32 mpad)
{
	const u32 r = nrows + j;
	const u32 off = S.inc_off[r], n = S.inc_len[r];
	u64 *tb = k4_tbits(S, wl, mpad);
	if (n <= B200_SHORT && wl <= 4) {
		// short row, narrow matrix (the usual case): list and column numbers by independent loads, the row's
		// words assembled in registers and stored once
		u32 l[B200_SHORT], col[B200_SHORT];
		load_short(S.inc_pool + off, n, l);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 q = 0; q < B200_SHORT; q++) col[q] = (q < n && l[q] != f) ? S.facet_local[l[q]] : B200_NONE;
		u64 acc[4] = {0, 0, 0, 0};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 q = 0; q < B200_SHORT; q++)
			if (col[q] != B200_NONE) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
				for (u32 w = 0; w < 4; w++)
					if ((col[q] >> 6) == w) acc[w] |= (u64)1 << (col[q] & 63);
				B200_ATOMIC_OR64(&tb[(size_t)col[q] * (mpad / 64) + (j >> 6)], (u64)1 << (j & 63));   // zeroed by k4_zero_cols
			}
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 w = 0; w < 4; w++)
			if (w < wl) S.bits[(size_t)w * mpad + j] = acc[w];
		return;
	}
	for (u32 w = 0; w < wl; w++) S.bits[(size_t)w * mpad + j] = 0;
	const u32 *l = S.inc_pool + off;
	for (u32 q = 0; q < n; q++) {
		const u32 fc = l[q];
		if (fc == f) continue;
		const u32 col = S.facet_local[fc];
		S.bits[(size_t)(col >> 6) * mpad + j] |= (u64)1 << (col & 63);
		B200_ATOMIC_OR64(&tb[(size_t)col * (mpad / 64) + (j >> 6)], (u64)1 << (j & 63));   // zeroed by k4_zero_cols
	}
}
B200_HD void k4_build_row(const DevState &S, u32 j)
{
	const CutCtl *c = S.ctl;
	k4_build_row_at(S, j, c->nrows, S.cur->facet, c->wl, c->mpad);
}
// the column matrix is filled with atomics, so it is cleared first (one barrier / launch earlier)
B200_HD void k4_zero_cols(const DevState &S, u64 x)
{
	k4_tbits(S, S.ctl->wl, S.ctl->mpad)[x] = 0;
}
// Containment by columns: a third row containing inc(a) & inc(b) exists iff the AND of the columns of
// all mask facets has a bit other than a and b (edge_test, bslv_poly.c:487-505).  Scalar form.
B200_HD bool k4_adjacent_by_columns(const DevState &S, u32 a, u32 b, u32 M, u32 wl, u32 mpad)
{
	if (S.d == 1) return true;
	const u64 *tb = k4_tbits(S, wl, mpad);
	const u32 mw = mpad / 64;
	for (u32 xw = 0; xw < mw; xw++) {
		u64 acc = (xw + 1) * 64 <= M ? ~(u64)0 : (M > xw * 64 ? (((u64)1 << (M - xw * 64)) - 1) : 0);   // valid rows only
		for (u32 w = 0; w < wl && acc; w++) {
			u64 m = S.bits[(size_t)w * mpad + a] & S.bits[(size_t)w * mpad + b];
			while (m && acc) {
				u32 bit = 0;
				while (!((m >> bit) & 1)) bit++;
				m &= m - 1;
				acc &= tb[(size_t)(w * 64 + bit) * mw + xw];
			}
		}
		if ((a >> 6) == xw) acc &= ~((u64)1 << (a & 63));
		if ((b >> 6) == xw) acc &= ~((u64)1 << (b & 63));
		if (acc) return false;
	}
	return true;
}
B200_HD void k4_push_survivor(const DevState &S, u32 a, u32 b)
{
	const u32 p = B200_ATOMIC_ADD(&S.ctl->n_surv, 1u);
	if (p < S.cap_pairs) { S.surv_a[p] = a; S.surv_b[p] = b; }
}
B200_HD void k4_push_pair(const DevState &S, u32 a, u32 b)
{
	const u32 p = B200_ATOMIC_ADD(&S.ctl->n_pairs, 1u);
	B200_ATOMIC_ADD(&S.deg[a], 1u);
	B200_ATOMIC_ADD(&S.deg[b], 1u);
	if (p < S.cap_pairs) { S.pair_a[p] = a; S.pair_b[p] = b; }
}
// scalar forms (host test double; the kernels use the tiled / warp-cooperative forms)
B200_HD void k4_filter_pair_in(const DevState &S, const u64 *bits, u32 wl, u32 mpad, u32 a, u32 b, u32 thr)
{
	u32 n = 0;
	for (u32 w = 0; w < wl; w++) n += popc64(bits[(size_t)w * mpad + a] & bits[(size_t)w * mpad + b]);
	if (n >= thr) k4_push_survivor(S, a, b);
}
// threshold of the popcount filter: d-1 common facets (bslv_poly.c:484); inside a cut the new facet is
// common to every row and left out of the matrix, hence d-2 there
B200_HD u32 k4_threshold(const DevState &S, bool new_facet_excluded)
{
	const u32 need = S.d >= 1 ? (u32)S.d - 1 : 0;
	return new_facet_excluded ? (need ? need - 1 : 0) : need;
}
B200_HD void k4_filter_pair(const DevState &S, u32 a, u32 b) { k4_filter_pair_in(S, S.bits, S.ctl->wl, S.ctl->mpad, a, b, k4_threshold(S, true)); }

// K6: the same pair test on the DUAL polytope -- rows = live facets, columns = live
// vertices (poly__update_adjacence(&dual), bslv_poly.c:992-1010 with edge_test on the dual).
B200_HD void k6_set_row_bits(const DevState &S, u32 r, u32 mpad)
{
	if (!bit_test(S.live, r)) return;
	const u32 *l = S.inc_pool + S.inc_off[r];
	for (u32 q = 0, n = S.inc_len[r]; q < n; q++) {
		const u32 col = S.facet_local[l[q]];
		if (col != B200_NONE) B200_ATOMIC_OR64(&S.bits[(size_t)(r >> 6) * mpad + col], (u64)1 << (r & 63));
	}
}
B200_HD void k4_contain_pair(const DevState &S, u32 s)
{
	const CutCtl *c = S.ctl;
	const u32 a = S.surv_a[s], b = S.surv_b[s];
	if (k4_adjacent_by_columns(S, a, b, c->n_new, c->wl, c->mpad)) k4_push_pair(S, a, b);
}
// row-scan form (K6: the dual polytope has 10^6 columns, its column matrix is not built)
B200_HD void k6_contain_pair(const DevState &S, u32 s)
{
	const CutCtl *c = S.ctl;
	const u32 a = S.surv_a[s], b = S.surv_b[s], M = c->n_new;
	bool adjacent = true;
	if (S.d != 1)
		for (u32 x = 0; x < M && adjacent; x++) {
			if (x == a || x == b) continue;
			bool contains = true;
			for (u32 w = 0; w < c->wl && contains; w++) {
				const u64 m = S.bits[(size_t)w * c->mpad + a] & S.bits[(size_t)w * c->mpad + b];
				contains = (S.bits[(size_t)w * c->mpad + x] & m) == m;
			}
			if (contains) adjacent = false;
		}
	if (adjacent) k4_push_pair(S, a, b);
}

// adjacency build: PLUS neighbours first, then the new-facet neighbours in ascending row order
B200_HD void adj_place(const DevState &S, u32 j)
{
	const u32 nw = S.ctl->nrows + j;
	const u32 off = S.ctl->adj_used + S.adj_base[j];
	const u32 np = S.new_padj_len[j], po = S.new_padj_off[j];
	S.adj_off[nw] = off;
	S.adj_len[nw] = np + S.deg[j];
	for (u32 q = 0; q < np; q++) S.adj_pool[off + q] = S.padj[po + q];
}
// adj_fill[j] counts the new-facet neighbours placed so far (zeroed when row j is emitted); the slot behind the
// PLUS neighbours is derived from the scan, so this stage does not wait for adj_place
B200_HD void adj_pair_fill(const DevState &S, u32 p)
{
	const u32 nrows = S.ctl->nrows, used = S.ctl->adj_used, a = S.pair_a[p], b = S.pair_b[p];
	const u32 oa = used + S.adj_base[a] + S.new_padj_len[a], ob = used + S.adj_base[b] + S.new_padj_len[b];
	const u32 pa = B200_ATOMIC_ADD(&S.adj_fill[a], 1u), pb = B200_ATOMIC_ADD(&S.adj_fill[b], 1u);
	S.adj_pool[oa + pa] = nrows + b;
	S.adj_pool[ob + pb] = nrows + a;
}
// rank-sort of up to N entries in registers (entries are distinct): one load and one store each
template <int N> B200_HD void adj_sort_regs(u32 *l, u32 m)
{
	u32 buf[N];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 x = 0; x < N; x++) buf[x] = x < m ? l[x] : B200_NONE;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (u32 x = 0; x < N; x++) {
		if (x >= m) break;
		u32 rk = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
		for (u32 y = 0; y < N; y++) rk += (buf[y] < buf[x]);
		l[rk] = buf[x];
	}
}
B200_HD void adj_sort(const DevState &S, u32 j)
{
	const u32 nw = S.ctl->nrows + j;
	const u32 lo = S.new_padj_len[j], n = S.adj_len[nw];
	u32 *l = S.adj_pool + S.adj_off[nw] + lo;
	const u32 m = n - lo;
	if (m <= 8) { adj_sort_regs<8>(l, m); return; }       // simple vertices: d-1 new-facet neighbours
	if (m <= 24) { adj_sort_regs<24>(l, m); return; }
	for (u32 x = 1; x < m; x++) {
		u32 key = l[x], y = x;
		while (y > 0 && l[y - 1] > key) { l[y] = l[y - 1]; y--; }
		l[y] = key;
	}
}

// ---- delta record for the host mirror (SURVEY 8(b) coherence rule), packed so that one D2H copy
// brings everything poly__add_vrtx has to apply.  Layout after the B200_STAGE_HDR-byte header:
//   double coords[n_new][d] | u32 parent_slot[n_new] | u32 dead_slot[n_vis] | u32 dead_facet[n_dead] | u8 ideal[n_new]
struct StageLayout { u64 coords, parent, dead_slots, dead_facets, ideal, total; };
B200_HD StageLayout stage_layout(const CutCtl &c, int d)
{
	StageLayout L;
	const bool cutting = !(c.status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE | ST_NEED_BIG));
	const u64 n_new = cutting ? c.n_new : 0, n_vis = cutting ? c.n_vis : 0, n_dead = cutting ? c.n_dead_facets : 0;
	L.coords = B200_STAGE_HDR;
	L.parent = L.coords + 8 * n_new * (u64)d;
	L.dead_slots = L.parent + 4 * n_new;
	L.dead_facets = L.dead_slots + 4 * n_vis;
	L.ideal = L.dead_facets + 4 * n_dead;
	L.total = (L.ideal + n_new + 7) & ~(u64)7;
	return L;
}
// item e of the packing pass; n_items = n_new*d + n_new + n_vis + n_dead  (ideal rides with parent)
B200_HD void pack_delta_item(const DevState &S, const StageLayout &L, u64 e, u32 first_row)
{
	const CutCtl *c = S.ctl;
	const u64 n_new = c->n_new, d = (u64)S.d;
	if (e < n_new * d) {
		const u64 r = e / d, j = e % d;
		((double *)(S.stage + L.coords))[e] = S.coord[j * S.cap_rows + first_row + r];
		return;
	}
	e -= n_new * d;
	if (e < n_new) {
		((u32 *)(S.stage + L.parent))[e] = S.new_parent[e];
		S.stage[L.ideal + e] = bit_test(S.ideal, first_row + (u32)e) ? 1 : 0;
		return;
	}
	e -= n_new;
	if (e < c->n_vis) { ((u32 *)(S.stage + L.dead_slots))[e] = S.dead_slots[e]; return; }
	e -= c->n_vis;
	((u32 *)(S.stage + L.dead_facets))[e] = S.dead_facets[e];
}
