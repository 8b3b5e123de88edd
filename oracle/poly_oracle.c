/*
 * TEST INFRASTRUCTURE ONLY -- see poly_oracle.h.
 *
 * CPU restatement of bensolve's double-description cut (reference: bslv_poly.c, BENSOLVE 2.0.1).
 * It is NOT a transcription: the reference walks the adjacency graph recursively and edits
 * unsorted index lists as it goes (bslv_poly.c:562-709); this file states the same mathematics
 * as the order-independent set semantics the CUDA engine implements:
 *
 *   1. trigger   : lowest live slot with h.x < thr - 1e-9            (bslv_poly.c:121-128)
 *   2. classes   : PLUS / ZERO+ / ZERO / MINUS per thresholds         (:573, :596, :666)
 *   3. visited   : vertices reachable from the trigger through non-PLUS vertices (:590-694)
 *   4. new slots : one per (MINUS, PLUS) edge (:597-627), one copy per ZERO vertex (:573-588)
 *   5. incidence : {f} u (inc(k) n inc(v)) resp. the union over PLUS neighbours (:634-665)
 *   6. adjacency : PLUS neighbours rewired (:628-633), then all pairs on the new facet through
 *                  the combinatorial test (:138-143, :467-512)
 *   7. facets    : a facet is dead iff it holds no live vertex (deliberate deviation: the
 *                  reference leaves order-dependent "ghost" facets, :697-705)
 *
 * Floating point follows the reference operation by operation (strict left-to-right sums,
 * separately rounded multiply and add; build with -ffp-contract=off), so coordinates come out
 * bit-identical to the reference, not merely within 1e-9.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "poly_oracle.h"

#define POLY_EPS 1e-9                          /* bslv_poly.h:47 */
#define SLOT_BITS 64
enum { C_UNSEEN = 0, C_PLUS, C_ZERO, C_MINUS };

static size_t g_default_dim;                   /* fnc_dim, bslv_poly.c:28 */
static oracle_cut_stats g_stats;
const oracle_cut_stats *oracle_last_cut_stats(void) { return &g_stats; }

/* ---------------------------------------------------------------- bitsets and lists */
static int bit_get(const size_t *w, size_t i) { return (int)((w[i / SLOT_BITS] >> (i % SLOT_BITS)) & 1u); }
static void bit_set(size_t *w, size_t i) { w[i / SLOT_BITS] |= (size_t)1 << (i % SLOT_BITS); }
static void bit_clr(size_t *w, size_t i) { w[i / SLOT_BITS] &= ~((size_t)1 << (i % SLOT_BITS)); }

static void list_init(poly_list *l) { l->cnt = 0; l->blcks = 0; l->data = NULL; }
static void list_push(poly_list *l, size_t x)
{
	if (l->cnt == l->blcks) {
		l->blcks = l->blcks ? 2 * l->blcks : 4;
		l->data = (size_t *)realloc(l->data, l->blcks * sizeof(size_t));
	}
	l->data[l->cnt++] = x;
}
static int list_has(const poly_list *l, size_t x)
{
	for (size_t i = 0; i < l->cnt; i++)
		if (l->data[i] == x) return 1;
	return 0;
}
static void list_drop(poly_list *l, size_t x)        /* remove one occurrence, order not kept */
{
	for (size_t i = 0; i < l->cnt; i++)
		if (l->data[i] == x) { l->data[i] = l->data[--l->cnt]; return; }
}
static void list_replace(poly_list *l, size_t from, size_t to)
{
	for (size_t i = 0; i < l->cnt; i++)
		if (l->data[i] == from) { l->data[i] = to; return; }
}

/* ---------------------------------------------------------------- slot storage */
static void poly_alloc(polytope *p)
{
	size_t cap = SLOT_BITS;
	p->cnt = 0;
	p->blcks = 1;
	p->ip = NULL;
	p->data = (double *)malloc(cap * (p->dim ? p->dim : 1) * sizeof(double));
	p->data_primg = (double *)malloc(cap * (p->dim_primg ? p->dim_primg : 1) * sizeof(double));
	p->adjacence = (poly_list *)malloc(cap * sizeof(poly_list));
	p->incidence = (poly_list *)malloc(cap * sizeof(poly_list));
	for (size_t i = 0; i < cap; i++) { list_init(p->adjacence + i); list_init(p->incidence + i); }
	p->used = (size_t *)calloc(1, sizeof(size_t));
	p->ideal = (size_t *)calloc(1, sizeof(size_t));
	p->sltn = (size_t *)calloc(1, sizeof(size_t));
}

/* append one live slot; returns its index (the reference's add_vrtx, bslv_poly.c:416-447) */
static size_t slot_append(polytope *p)
{
	size_t cap = p->blcks * SLOT_BITS;
	if (p->cnt + 1 >= cap) {
		size_t nb = 2 * p->blcks, ncap = nb * SLOT_BITS;
		p->data = (double *)realloc(p->data, ncap * (p->dim ? p->dim : 1) * sizeof(double));
		p->data_primg = (double *)realloc(p->data_primg, ncap * (p->dim_primg ? p->dim_primg : 1) * sizeof(double));
		p->adjacence = (poly_list *)realloc(p->adjacence, ncap * sizeof(poly_list));
		p->incidence = (poly_list *)realloc(p->incidence, ncap * sizeof(poly_list));
		for (size_t i = cap; i < ncap; i++) { list_init(p->adjacence + i); list_init(p->incidence + i); }
		p->used = (size_t *)realloc(p->used, nb * sizeof(size_t));
		p->ideal = (size_t *)realloc(p->ideal, nb * sizeof(size_t));
		p->sltn = (size_t *)realloc(p->sltn, nb * sizeof(size_t));
		for (size_t i = p->blcks; i < nb; i++) p->used[i] = p->ideal[i] = p->sltn[i] = 0;
		p->blcks = nb;
	}
	bit_set(p->used, p->cnt);
	return p->cnt++;
}

static void poly_free(polytope *p)
{
	for (size_t i = 0; i < p->blcks * SLOT_BITS; i++) { free(p->adjacence[i].data); free(p->incidence[i].data); }
	free(p->adjacence); free(p->incidence);
	free(p->data); free(p->data_primg); free(p->used); free(p->ideal); free(p->sltn);
}

/* ---------------------------------------------------------------- API: setup */
static void default_dual_to_halfspace(double *dual_point, int is_dir, double *hp)
{	/* cone_polar, bslv_poly.c:30-39: d -> { y : d.y >= -1 }  (>= 0 for a direction) */
	for (size_t j = 0; j < g_default_dim; j++) hp[j] = dual_point[j];
	hp[g_default_dim] = is_dir ? 0 : -1.0;
}

void poly__set_default_args(poly_args *a, size_t dim)
{
	g_default_dim = dim;
	a->dim = dim;
	a->eps = 1e-08;                       /* written, never read (bslv_poly.c:45) */
	a->dim_primg_prml = a->dim_primg_dl = 0;
	a->primalV2dualH = NULL;
	a->dualV2primalH = default_dual_to_halfspace;
}

void poly__initialise(poly_args *a)
{
	size_t d = a->dim;
	a->primal.dim = a->dual.dim = d;
	a->primal.dim_primg = a->dim_primg_prml;
	a->dual.dim_primg = a->dim_primg_dl;
	poly_alloc(&a->primal);
	poly_alloc(&a->dual);
	a->primal.dual = &a->dual;
	a->dual.dual = &a->primal;
	a->primal.v2h = a->primalV2dualH;
	a->dual.v2h = a->dualV2primalH;
	a->val = (double *)malloc((d ? d : 1) * sizeof(double));
	a->val_primg_prml = (double *)malloc((a->dim_primg_prml ? a->dim_primg_prml : 1) * sizeof(double));
	a->val_primg_dl = (double *)malloc((a->dim_primg_dl ? a->dim_primg_dl : 1) * sizeof(double));
	/* dual slot 0 = (0,...,0,-1), ideal: the facet at infinity (bslv_poly.c:83-92) */
	size_t f0 = slot_append(&a->dual);
	for (size_t j = 0; j < d; j++) a->dual.data[f0 * d + j] = (j + 1 == d) ? -1.0 : 0.0;
	for (size_t j = 0; j < a->dim_primg_dl; j++) a->dual.data_primg[j] = 0;
	bit_set(a->dual.ideal, f0);
	a->init_data.H = (double *)malloc(d * d * sizeof(double));
	a->init_data.R = (double *)malloc(d * (d + 1) / 2 * sizeof(double));
	a->init_data.alph = (double *)malloc(d * sizeof(double));
	list_init(&a->init_data.queue);
	list_init(&a->init_data.gnrtrs);
	a->init_data.intlsd = 0;
}

void poly__kill(poly_args *a)
{
	poly_free(&a->primal);
	poly_free(&a->dual);
	free(a->val); free(a->val_primg_prml); free(a->val_primg_dl);
	if (!a->init_data.intlsd) {
		free(a->init_data.H); free(a->init_data.R); free(a->init_data.alph);
		free(a->init_data.queue.data); free(a->init_data.gnrtrs.data);
	}
}

int poly__get_vrtx(poly_args *a)
{	/* first slot that is live and not yet marked as solution (bslv_poly.c:210-226) */
	const polytope *p = &a->primal;
	size_t s = 0;
	while (s < p->cnt && !(bit_get(p->used, s) && !bit_get(p->sltn, s))) s++;
	a->idx = s;
	if (s == p->cnt) return EXIT_FAILURE;
	a->ideal = (unsigned)bit_get(p->ideal, s);
	for (size_t j = 0; j < p->dim; j++) a->val[j] = p->data[s * p->dim + j];
	return EXIT_SUCCESS;
}

/* ---------------------------------------------------------------- combinatorial adjacency */
/* u ~ w  <=>  |inc(u) n inc(w)| >= d-1 and no third vertex contains that intersection
 * (bslv_poly.c:467-512; d == 1 => always adjacent).  Any such third vertex lies on every mutual
 * facet, so scanning the vertex list of one of them is exhaustive. */
static int adjacent_by_incidence(const polytope *p, size_t u, size_t w)
{
	const poly_list *iu = p->incidence + u, *iw = p->incidence + w;
	size_t nm = 0, *mutual = (size_t *)malloc((iu->cnt ? iu->cnt : 1) * sizeof(size_t));
	int ok = 1;
	for (size_t i = 0; i < iu->cnt; i++)
		if (list_has(iw, iu->data[i])) mutual[nm++] = iu->data[i];
	if (p->dim == 1) ok = 1;
	else if (nm + 1 < p->dim) ok = 0;
	else {
		const poly_list *cand = p->dual->incidence + mutual[0];
		for (size_t c = 0; c < cand->cnt && ok; c++) {
			size_t x = cand->data[c];
			if (x == u || x == w) continue;
			size_t m = 1;
			while (m < nm && list_has(p->incidence + x, mutual[m])) m++;
			if (m == nm) ok = 0;
		}
	}
	free(mutual);
	return ok;
}

void poly__update_adjacence(polytope *p)
{	/* all pairs of live slots (bslv_poly.c:992-1010); appends, as the reference does */
	for (size_t u = 0; u < p->cnt; u++) {
		if (!bit_get(p->used, u)) continue;
		for (size_t w = u + 1; w < p->cnt; w++)
			if (bit_get(p->used, w) && adjacent_by_incidence(p, u, w)) {
				list_push(p->adjacence + u, w);
				list_push(p->adjacence + w, u);
			}
	}
}

/* ---------------------------------------------------------------- the cut */
static double dot_lr(const double *h, const double *x, size_t d)
{	/* strict left-to-right, starting from 0 (bslv_poly.c:123-125, 569-571, 593-595) */
	double s = 0;
	for (size_t j = 0; j < d; j++) s += h[j] * x[j];
	return s;
}

static int cut_with_halfspace(poly_args *a, size_t f, const double *hp)
{
	polytope *P = &a->primal, *D = &a->dual;
	const size_t d = a->dim, S = P->cnt;
	const double alpha = hp[d];
	memset(&g_stats, 0, sizeof g_stats);

	/* 1. trigger scan */
	size_t v0 = S;
	for (size_t s = 0; s < S; s++) {
		if (!bit_get(P->used, s)) continue;
		double t = dot_lr(hp, P->data + s * d, d);
		if (t < (bit_get(P->ideal, s) ? 0 : alpha) - POLY_EPS) { v0 = s; break; }
	}
	a->idx = v0;
	if (v0 == S) { bit_clr(D->used, f); return EXIT_FAILURE; }      /* redundant (:132-136) */

	/* 2+3. classes of everything reachable from v0 through non-PLUS vertices */
	unsigned char *cls = (unsigned char *)calloc(S, 1);
	size_t *visited = (size_t *)malloc(S * sizeof(size_t)), nvis = 0, head = 0;
	cls[v0] = C_MINUS;
	visited[nvis++] = v0;
	while (head < nvis) {
		size_t v = visited[head++];
		const poly_list *av = P->adjacence + v;
		for (size_t i = 0; i < av->cnt; i++) {
			size_t k = av->data[i];
			if (cls[k] != C_UNSEEN) continue;
			double *x = P->data + k * d;
			double thr = bit_get(P->ideal, k) ? 0 : alpha;
			double t = dot_lr(hp, x, d);
			if (t > thr + POLY_EPS) { cls[k] = C_PLUS; continue; }
			if (t > thr + 1.0e-2 * POLY_EPS) {                          /* ZERO+: project (:666-673) */
				double mu = t - thr, hh = 0;
				for (size_t j = 0; j < d; j++) hh += hp[j] * hp[j];
				mu /= hh;
				for (size_t j = 0; j < d; j++) x[j] -= mu * hp[j];
				t = dot_lr(hp, x, d);
				g_stats.n_zero_plus_projected++;
			}
			cls[k] = (t > thr - POLY_EPS) ? C_ZERO : C_MINUS;           /* (:573) */
			visited[nvis++] = k;
		}
	}
	/* convexity check for the tests: is every non-PLUS live vertex reached? */
	for (size_t s = 0; s < S; s++) {
		if (!bit_get(P->used, s) || cls[s] != C_UNSEEN) continue;
		double t = dot_lr(hp, P->data + s * d, d);
		if (!(t > (bit_get(P->ideal, s) ? 0 : alpha) + 1.0e-2 * POLY_EPS)) g_stats.n_nonplus_unreached++;
	}
	/* deterministic numbering of the new slots: visited vertices by ascending slot */
	for (size_t i = 1; i < nvis; i++) {
		size_t x = visited[i], j = i;
		while (j && visited[j - 1] > x) { visited[j] = visited[j - 1]; j--; }
		visited[j] = x;
	}

	/* 4-6a. new slots */
	double *dir = (double *)malloc(d * sizeof(double));
	for (size_t vi = 0; vi < nvis; vi++) {
		const size_t v = visited[vi];
		const int v_ideal = bit_get(P->ideal, v);
		size_t copy = (size_t)-1;
		if (cls[v] == C_ZERO) {                                           /* copy (:573-588) */
			copy = slot_append(P);
			memcpy(P->data + copy * d, P->data + v * d, d * sizeof(double));
			if (v_ideal) bit_set(P->ideal, copy);
			if (bit_get(P->sltn, v)) {
				bit_set(P->sltn, copy);
				memcpy(P->data_primg + copy * P->dim_primg, P->data_primg + v * P->dim_primg,
				       P->dim_primg * sizeof(double));
			}
			list_push(P->incidence + copy, f);
			list_push(D->incidence + f, copy);
			g_stats.n_copies++;
		} else
			g_stats.n_minus++;
		for (size_t i = 0; i < P->adjacence[v].cnt; i++) {
			const size_t k = P->adjacence[v].data[i];
			if (k >= S || cls[k] != C_PLUS) continue;
			const int k_ideal = bit_get(P->ideal, k);
			size_t nv = copy;
			if (cls[v] == C_MINUS) {                                      /* edge vertex (:597-627) */
				nv = slot_append(P);
				double *out = P->data + nv * d;
				const double *xv = P->data + v * d, *xk = P->data + k * d;
				const double *base = k_ideal ? xv : xk;
				const double *dsrc = k_ideal ? xk : xv;
				for (size_t j = 0; j < d; j++) { out[j] = base[j]; dir[j] = dsrc[j]; }
				double rhs = alpha;
				if (k_ideal && v_ideal) {
					for (size_t j = 0; j < d; j++) dir[j] -= xv[j];
					bit_set(P->ideal, nv);
					rhs = 0;
				} else if (!k_ideal && !v_ideal)
					for (size_t j = 0; j < d; j++) dir[j] -= xk[j];
				double mu = rhs - dot_lr(hp, out, d);
				mu /= dot_lr(hp, dir, d);
				for (size_t j = 0; j < d; j++) out[j] += mu * dir[j];
				list_push(P->incidence + nv, f);
				list_push(D->incidence + f, nv);
				g_stats.n_edge_vertices++;
			}
			/* adjacency: k now neighbours the new slot instead of v (:628-633) */
			list_replace(P->adjacence + k, v, nv);
			list_push(P->adjacence + nv, k);
			/* incidence: facets shared by v and k (:634-665) */
			for (size_t q = 0; q < P->incidence[k].cnt; q++) {
				size_t fc = P->incidence[k].data[q];
				if (!list_has(P->incidence + v, fc) || list_has(P->incidence + nv, fc)) continue;
				list_push(P->incidence + nv, fc);
				list_push(D->incidence + fc, nv);
			}
		}
	}
	free(dir);

	/* 7. delete visited vertices; a facet without live vertices dies */
	for (size_t vi = 0; vi < nvis; vi++) {
		size_t v = visited[vi];
		bit_clr(P->used, v);
		for (size_t q = 0; q < P->incidence[v].cnt; q++) list_drop(D->incidence + P->incidence[v].data[q], v);
	}
	for (size_t vi = 0; vi < nvis; vi++) {
		size_t v = visited[vi];
		for (size_t q = 0; q < P->incidence[v].cnt; q++) {
			size_t fc = P->incidence[v].data[q];
			if (D->incidence[fc].cnt == 0) bit_clr(D->used, fc);
		}
	}
	g_stats.n_zero = g_stats.n_copies;
	free(visited);
	free(cls);

	/* 6b. adjacency among the vertices of the new facet (:138-143) */
	const poly_list *nf = D->incidence + f;
	for (size_t i = 0; i < nf->cnt; i++)
		for (size_t j = 0; j < i; j++) {
			g_stats.n_pair_tests++;
			if (adjacent_by_incidence(P, nf->data[i], nf->data[j])) {
				list_push(P->adjacence + nf->data[i], nf->data[j]);
				list_push(P->adjacence + nf->data[j], nf->data[i]);
				g_stats.n_new_adjacent_pairs++;
			}
		}
	return EXIT_SUCCESS;
}

int poly__add_vrtx(poly_args *a)
{
	polytope *D = &a->dual;
	const size_t d = a->dim;
	size_t f = slot_append(D);                                           /* (:109-116) */
	if (a->ideal) bit_set(D->ideal, f);
	for (size_t j = 0; j < d; j++) D->data[f * d + j] = a->val[j];
	for (size_t j = 0; j < a->dim_primg_dl; j++) D->data_primg[f * D->dim_primg + j] = a->val_primg_dl[j];
	if (!a->init_data.intlsd) {                                          /* queue until init (:145) */
		list_push(&a->init_data.queue, f);
		return EXIT_SUCCESS;
	}
	double *hp = (double *)malloc((d + 1) * sizeof(double));
	a->dualV2primalH(a->val, (int)a->ideal, hp);                         /* (:119) */
	int rc = cut_with_halfspace(a, f, hp);
	free(hp);
	return rc;
}

/* ---------------------------------------------------------------- start simplex */
static double norm2(const double *x, size_t n)
{
	double s = 0;
	for (size_t l = 0; l < n; l++) s += x[l] * x[l];
	return sqrt(s);
}

/* One modified Gram-Schmidt step (bslv__normalise, bslv_poly.c:1030-1060): orthonormalise x
 * against rows 0..k-1 of Q into row k, write row k of the packed lower-triangular R, return the
 * relative residual (0 if the residual norm is below 1e-6). */
static double gram_schmidt_step(const double *x, double *Q, double *R, size_t k, size_t n)
{
	double nrm_in = norm2(x, n), *qk = Q + k * n;
	memcpy(qk, x, n * sizeof(double));
	for (size_t j = 0; j < k; j++) {
		double s = 0;
		for (size_t l = 0; l < n; l++) s += Q[j * n + l] * qk[l];
		for (size_t l = 0; l < n; l++) qk[l] -= s * Q[j * n + l];
	}
	double res = norm2(qk, n);
	if (res < 1.0e-6) return 0;
	for (size_t l = 0; l < n; l++) qk[l] /= res;
	for (size_t j = 0; j <= k; j++) {
		double s = 0;
		for (size_t l = 0; l < n; l++) s += Q[j * n + l] * x[l];
		R[k * (k + 1) / 2 + j] = s;
	}
	return res / nrm_in;
}

/* Start polyhedron {y : Q_k . y >= alph_k}: one vertex and d extreme directions
 * (poly__poly_initialise, bslv_poly.c:711-787). perm[0] is the facet at infinity. */
static void build_start_simplex(polytope *P, const double *Q, const double *R, const double *alph, const size_t *perm)
{
	const size_t d = P->dim;
	double *z = (double *)calloc(d, sizeof(double));         /* R z = alph (forward substitution) */
	double *T = (double *)calloc(d * d, sizeof(double));     /* T[l][k]: column k solves against e_k */
#define RR(k, j) R[(k) * ((k) + 1) / 2 + (j)]
	for (size_t k = 0; k < d; k++) {
		z[k] = alph[k];
		T[k * d + k] = 1.0;
		for (size_t j = 0; j < k; j++) {
			z[k] -= z[j] * RR(k, j);
			for (size_t l = 0; l < d; l++) T[l * d + k] -= T[l * d + j] * RR(k, j);
		}
		z[k] /= RR(k, k);
		for (size_t l = 0; l < d; l++) T[l * d + k] /= RR(k, k);
	}
#undef RR
	size_t v = slot_append(P);                               /* slot 0: the vertex Q^T z */
	for (size_t k = 0; k < d; k++) {
		double s = 0;
		for (size_t j = 0; j < d; j++) s += z[j] * Q[j * d + k];
		P->data[v * d + k] = s;
	}
	for (size_t k = 0; k < d; k++) {                          /* slots 1..d: directions */
		size_t r = slot_append(P);
		bit_set(P->ideal, r);
		for (size_t j = 0; j < d; j++) {
			double s = 0;
			for (size_t l = 0; l < d; l++) s += Q[l * d + j] * T[k * d + l];
			P->data[r * d + j] = s;
		}
	}
	for (size_t k = 0; k <= d; k++)                           /* facet perm[k] holds all but vertex k */
		for (size_t j = 0; j <= d; j++)
			if (j != k) {
				list_push(P->dual->incidence + perm[k], j);
				list_push(P->incidence + j, perm[k]);
				list_push(P->adjacence + k, j);
			}
	free(z);
	free(T);
}

int poly__intl_apprx(poly_args *a)
{
	const size_t d = a->dim;
	poly_list *Qu = &a->init_data.queue, *G = &a->init_data.gnrtrs;
	if (Qu->cnt < d) return EXIT_FAILURE;                                /* (:158-159) */
	double *hp = (double *)malloc((d + 1) * Qu->cnt * sizeof(double));
	for (size_t q = 0; q < Qu->cnt; q++)
		a->dualV2primalH(a->dual.data + Qu->data[q] * d, bit_get(a->dual.ideal, Qu->data[q]), hp + q * (d + 1));
	size_t *perm = (size_t *)malloc((d + 1) * sizeof(size_t));
	perm[0] = 0;
	while (G->cnt < d) {                                                 /* greedy pivoting (:167-185) */
		double best = 0;
		size_t arg = 0;
		for (size_t q = 0; q < Qu->cnt; q++) {
			double r = gram_schmidt_step(hp + q * (d + 1), a->init_data.H, a->init_data.R, G->cnt, d);
			if (best < r) { best = r; arg = q; }
		}
		if (best < 1.0e-10) { free(hp); free(perm); return EXIT_FAILURE; }
		gram_schmidt_step(hp + arg * (d + 1), a->init_data.H, a->init_data.R, G->cnt, d);
		a->init_data.alph[G->cnt] = hp[arg * (d + 1) + d];
		list_push(G, Qu->data[arg]);
		perm[G->cnt] = Qu->data[arg];
		memmove(hp + arg * (d + 1), hp + (Qu->cnt - 1) * (d + 1), (d + 1) * sizeof(double));
		Qu->data[arg] = Qu->data[--Qu->cnt];
	}
	build_start_simplex(&a->primal, a->init_data.H, a->init_data.R, a->init_data.alph, perm);
	free(perm);
	a->init_data.intlsd = 1;
	/* the halfspaces still queued are retired and re-added as fresh dual slots (:190-197) */
	for (size_t q = 0; q < Qu->cnt; q++) bit_clr(a->dual.used, Qu->data[q]);
	for (size_t q = 0; q < Qu->cnt; q++) {
		size_t src = Qu->data[q];
		for (size_t j = 0; j < d; j++) a->val[j] = a->dual.data[src * d + j];
		a->ideal = (unsigned)bit_get(a->dual.ideal, src);
		poly__add_vrtx(a);
	}
	free(hp);
	free(Qu->data); free(G->data);
	free(a->init_data.H); free(a->init_data.R); free(a->init_data.alph);
	return EXIT_SUCCESS;
}
