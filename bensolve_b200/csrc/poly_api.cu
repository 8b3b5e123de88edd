// The reference-facing C ABI (include/bensolve_b200.h): same symbols, struct layouts, argument
// meaning and return codes as bslv_poly.h:90-118, so that bensolve's unchanged bslv_algs.c links
// against this library instead of bslv_poly.o.
//
// What runs where:
//   device : the cut (poly__add_vrtx after initialisation) -- cut_engine.cu / cut_kernels.cuh
//   host   : the AoS mirror the caller reads and writes between calls (data, used, ideal, cnt;
//            sltn and data_primg are host-authoritative), the whole dual polytope (one row per
//            halfspace), the O(d^3) start simplex, result writers, lazily materialised lists.
// CUDA errors have no channel in this API (void / int returns the caller mostly ignores), so they
// print a message and abort().
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>
#include <chrono>
#include <thread>
#include <mutex>
#include <sys/mman.h>
#include <vector>

#include "../../include/bensolve_b200.h"
#include "cut_engine.h"

#define SLOTS_PER_BLOCK 64
#define HANDLE_MAGIC 0xB200C0DEB200C0DEull

struct Handle {                 // hung off the dead `ip` field of both polytopes (bslv_poly.c:233)
	u64 magic = HANDLE_MAGIC;
	poly_args *owner = nullptr;
	CutEngine *engine = nullptr;
	bool lists_current = false;   // host poly_lists reflect device state
	bool host_may_have_edited = false; // caller edits primal.data in place at end of run (bslv_algs.c:193-273)
	size_t get_cursor = 0;        // every slot below is dead or already marked sltn (f2)
	std::vector<size_t> slab_pinc, slab_padj, slab_dinc, slab_dadj; // backing store of the host lists
	unsigned flags = 0;
	size_t lists_cap_p = 0, lists_cap_d = 0;   // slots the host poly_list arrays are sized for
	double prof_apply_us = 0, prof_total_us = 0;   // B200_PHASES accumulators (per polytope: several are alive at once)
	u64 prof_calls = 0;
};

static size_t g_default_dim; // fnc_dim, bslv_poly.c:28: read by the default callback

[[noreturn]] static void die(const char *where, const char *what)
{
	fprintf(stderr, "bensolve_b200: fatal error in %s: %s\n", where, what);
	abort();
}
#define GUARD_BEGIN try {
#define GUARD_END(where) } catch (const std::exception &e) { die(where, e.what()); }

static Handle *handle_of(const polytope *p)
{
	Handle *h = reinterpret_cast<Handle *>(p->ip);
	if (!h || h->magic != HANDLE_MAGIC) die("handle_of", "polytope was not created by poly__initialise of this library");
	return h;
}

// ------------------------------------------------------------------ host mirror storage
static void list_reset(poly_list *l) { l->cnt = 0; l->blcks = 0; l->data = NULL; }

// The coordinate array of a 10^6-vertex polytope is a few hundred MB.  Fresh memory of that size costs more in page
// faults and kernel zero-fill than the cuts that fill it, so blocks released by poly__kill are kept (a few, bounded;
// B200_HOST_CACHE_MB, default 1024, 0 = off) and handed to the next polytope of this process.  They are ordinary
// malloc blocks.
namespace {
struct BigCache {
	std::mutex mu;
	struct E { void *p; size_t bytes; };
	std::vector<E> held;
	size_t held_bytes = 0;
} g_big;
const size_t BIG_MIN = (size_t)1 << 20;
size_t big_limit()
{
	static const size_t lim = [] { const char *e = getenv("B200_HOST_CACHE_MB"); return (size_t)(e ? atol(e) : 1024) << 20; }();
	return lim;
}
void *big_take(size_t bytes)        // a recycled block of at least `bytes`, or NULL
{
	{
		std::lock_guard<std::mutex> lk(g_big.mu);
		size_t best = (size_t)-1;
		for (size_t i = 0; i < g_big.held.size(); i++)
			if (g_big.held[i].bytes >= bytes && g_big.held[i].bytes <= 4 * bytes && (best == (size_t)-1 || g_big.held[i].bytes < g_big.held[best].bytes)) best = i;
		if (best != (size_t)-1) {
			BigCache::E e = g_big.held[best];
			g_big.held.erase(g_big.held.begin() + best);
			g_big.held_bytes -= e.bytes;
			return e.p;
		}
	}
	return NULL;
}
void big_free(void *p, size_t bytes)
{
	if (!p) return;
	if (bytes >= BIG_MIN && big_limit()) {
		std::lock_guard<std::mutex> lk(g_big.mu);
		if (g_big.held.size() < 4 && g_big.held_bytes + bytes <= big_limit()) {
			g_big.held.push_back({p, bytes});
			g_big.held_bytes += bytes;
			return;
		}
	}
	free(p);
}
}   // namespace

// First touch of a fresh coordinate block costs ~0.5 ms per MB in page faults, on the thread that applies the delta
// records.  A helper thread pre-faults the newly grown tail with MADV_POPULATE_WRITE (contents untouched, so it is safe
// beside the writer) while the caller waits for the device anyway.  One request at a time; joined before the block
// moves or is freed.
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static std::thread &populate_thread() { static std::thread *t = new std::thread(); return *t; }   // (never destroyed: no join at exit)
static void populate_wait() { if (populate_thread().joinable()) populate_thread().join(); }
static void populate_async(char *from, size_t bytes)
{
	static const bool on = [] { const char *e = getenv("B200_HOST_POPULATE"); return !e || atoi(e) != 0; }();
	if (!on || bytes < ((size_t)1 << 22)) return;
	const uintptr_t lo = ((uintptr_t)from + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)from + bytes) & ~(uintptr_t)4095;
	if (hi <= lo) return;
	populate_thread() = std::thread([lo, hi] {
		for (uintptr_t a = lo; a < hi; a += (size_t)1 << 19)          // in small slices: the call holds the address-space lock cudaMalloc / cudaFree need
			if (madvise((void *)a, std::min<size_t>((size_t)1 << 19, hi - a), MADV_POPULATE_WRITE) != 0) break;
	});
}

static void mirror_alloc(polytope *p)
{
	const size_t cap = SLOTS_PER_BLOCK;
	p->cnt = 0;
	p->blcks = 1;
	p->data = (double *)malloc(cap * std::max<size_t>(p->dim, 1) * sizeof(double));
	p->data_primg = (double *)malloc(cap * std::max<size_t>(p->dim_primg, 1) * sizeof(double));
	p->adjacence = NULL;                  // the list arrays are allocated when they are first materialised
	p->incidence = NULL;
	p->used = (vrtx_strg *)calloc(1, sizeof(vrtx_strg));
	p->ideal = (vrtx_strg *)calloc(1, sizeof(vrtx_strg));
	p->sltn = (vrtx_strg *)calloc(1, sizeof(vrtx_strg));
}

static double g_mirror_grow_us = 0;       // B200_PHASES report only (process-wide, unsynchronised)
struct MirrorGrowTimer {
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	~MirrorGrowTimer() { g_mirror_grow_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
};
static void mirror_reserve(polytope *p, size_t slots)
{
	size_t cap = p->blcks * SLOTS_PER_BLOCK;
	if (slots < cap) return;              // keep one spare slot, as the reference does (bslv_poly.c:418)
	MirrorGrowTimer mgt;
	size_t nb = p->blcks;
	while (nb * SLOTS_PER_BLOCK <= slots) nb *= 2;
	const size_t ncap = nb * SLOTS_PER_BLOCK;
	const size_t row = std::max<size_t>(p->dim, 1) * sizeof(double), old_bytes = cap * row, new_bytes = ncap * row;
	if (new_bytes >= BIG_MIN) {           // large: a recycled block (see big_alloc) if one fits ...
		populate_wait();
		double *nd = (double *)big_take(new_bytes);
		if (nd) {
			memcpy(nd, p->data, p->cnt * row);
			big_free(p->data, old_bytes);
			p->data = nd;
		} else {                          // ... else grow in place: glibc moves the pages of a large block (mremap), no copy,
			p->data = (double *)realloc(p->data, new_bytes);      // no second round of first-touch faults
			if (!p->data) die("mirror_reserve", "out of host memory");
			populate_async((char *)p->data + old_bytes, new_bytes - old_bytes);
		}
	} else
		p->data = (double *)realloc(p->data, new_bytes);
	p->data_primg = (double *)realloc(p->data_primg, ncap * std::max<size_t>(p->dim_primg, 1) * sizeof(double));
	p->used = (vrtx_strg *)realloc(p->used, nb * sizeof(vrtx_strg));
	p->ideal = (vrtx_strg *)realloc(p->ideal, nb * sizeof(vrtx_strg));
	p->sltn = (vrtx_strg *)realloc(p->sltn, nb * sizeof(vrtx_strg));
	for (size_t i = p->blcks; i < nb; i++) p->used[i] = p->ideal[i] = p->sltn[i] = 0;
	p->blcks = nb;
}

// the poly_list arrays are only read inside this library (writers, polyck, update_adjacence, swap,
// plot -- bslv_algs.c never touches them), so they are (re)sized when lists are materialised
static void mirror_lists_reserve(polytope *p, Handle *h, bool is_primal)
{
	size_t &have = is_primal ? h->lists_cap_p : h->lists_cap_d;
	const size_t need = p->blcks * SLOTS_PER_BLOCK;
	if (have >= need && p->adjacence) return;
	p->adjacence = (poly_list *)realloc(p->adjacence, need * sizeof(poly_list));
	p->incidence = (poly_list *)realloc(p->incidence, need * sizeof(poly_list));
	for (size_t i = have; i < need; i++) { list_reset(p->adjacence + i); list_reset(p->incidence + i); }
	have = need;
}

static size_t mirror_append(polytope *p)      // add_vrtx, bslv_poly.c:416-447
{
	mirror_reserve(p, p->cnt + 1);
	ST_BT(p->used, p->cnt);
	return p->cnt++;
}

static void mirror_free(polytope *p)
{
	populate_wait();
	big_free(p->data, p->blcks * SLOTS_PER_BLOCK * std::max<size_t>(p->dim, 1) * sizeof(double));
	free(p->data_primg); free(p->adjacence); free(p->incidence);
	free(p->used); free(p->ideal); free(p->sltn);
}

// ------------------------------------------------------------------ defaults and life cycle
static void default_dual_to_halfspace(double *dual_point, int is_dir, double *hp)
{	// cone_polar (bslv_poly.c:30-39): dual point d  ->  halfspace d.y >= -1  (>= 0 for a direction)
	for (size_t j = 0; j < g_default_dim; j++) hp[j] = dual_point[j];
	hp[g_default_dim] = is_dir ? 0 : -1.0;
}

extern "C" void poly__set_default_args(poly_args *args, size_t dim)
{
	g_default_dim = dim;
	args->dim = dim;
	args->eps = 1e-08;
	args->dim_primg_prml = 0;
	args->dim_primg_dl = 0;
	args->primalV2dualH = NULL;
	args->dualV2primalH = (void (*)())default_dual_to_halfspace;
}

extern "C" void poly__initialise(poly_args *args)
{
	GUARD_BEGIN
	const size_t d = args->dim;
	if (d < 1 || d > B200_MAXD) die("poly__initialise", "dimension outside 1..16");
	args->primal.dim = args->dual.dim = d;
	args->primal.dim_primg = args->dim_primg_prml;
	args->dual.dim_primg = args->dim_primg_dl;
	mirror_alloc(&args->primal);
	mirror_alloc(&args->dual);
	args->primal.dual = &args->dual;
	args->dual.dual = &args->primal;
	args->primal.v2h = (void (*)(double *, int, double *))args->primalV2dualH;
	args->dual.v2h = (void (*)(double *, int, double *))args->dualV2primalH;
	Handle *h = new Handle();
	h->owner = args;
	args->primal.ip = reinterpret_cast<double *>(h);
	args->dual.ip = reinterpret_cast<double *>(h);
	args->val = (double *)malloc(std::max<size_t>(d, 1) * sizeof(double));
	args->val_primg_prml = (double *)malloc(std::max<size_t>(args->dim_primg_prml, 1) * sizeof(double));
	args->val_primg_dl = (double *)malloc(std::max<size_t>(args->dim_primg_dl, 1) * sizeof(double));
	// dual slot 0 = (0,...,0,-1) flagged ideal: the facet at infinity (bslv_poly.c:83-92)
	const size_t f0 = mirror_append(&args->dual);
	for (size_t j = 0; j < d; j++) args->dual.data[f0 * d + j] = (j + 1 == d) ? -1.0 : 0.0;
	for (size_t j = 0; j < args->dim_primg_dl; j++) args->dual.data_primg[j] = 0;
	ST_BT(args->dual.ideal, f0);
	args->init_data.H = (double *)malloc(d * d * sizeof(double));
	args->init_data.R = (double *)malloc(d * (d + 1) / 2 * sizeof(double));
	args->init_data.alph = (double *)malloc(d * sizeof(double));
	list_reset(&args->init_data.queue);
	list_reset(&args->init_data.gnrtrs);
	args->init_data.intlsd = 0;
	GUARD_END("poly__initialise")
}

extern "C" void poly__kill(poly_args *args)
{
	Handle *h = handle_of(&args->primal);
	if (getenv("B200_PHASES")) fprintf(stderr, "[b200] host mirror growth so far in this process: %.1f ms\n", g_mirror_grow_us / 1e3);
	delete h->engine;
	h->magic = 0;
	delete h;
	mirror_free(&args->primal);
	mirror_free(&args->dual);
	free(args->val); free(args->val_primg_prml); free(args->val_primg_dl);
	if (!args->init_data.intlsd) {
		free(args->init_data.H); free(args->init_data.R); free(args->init_data.alph);
		free(args->init_data.queue.data); free(args->init_data.gnrtrs.data);
	}
	args->primal.ip = args->dual.ip = NULL;
}

static void queue_push(poly_list *l, size_t x)
{
	if (l->cnt == l->blcks) {
		l->blcks = l->blcks ? 2 * l->blcks : 8;
		l->data = (size_t *)realloc(l->data, l->blcks * sizeof(size_t));
	}
	l->data[l->cnt++] = x;
}

// ------------------------------------------------------------------ the cut entry point
static void make_params(const double *hp, size_t d, u32 facet, CutParams &P)
{
	memset(&P, 0, sizeof P);
	double hh = 0;
	for (size_t j = 0; j < d; j++) { P.h[j] = hp[j]; hh += hp[j] * hp[j]; }
	P.alpha = hp[d];
	for (int id = 0; id < 2; id++) {          // thresholds exactly as bslv_poly.c:596, :666, :573/:126
		const double thr = id ? 0 : P.alpha;
		P.hi[id] = thr + POLY_EPS;
		P.mid[id] = thr + 1.0e-2 * POLY_EPS;
		P.lo[id] = thr - POLY_EPS;
	}
	P.hh = hh;
	P.facet = facet;
}

static void apply_delta(poly_args *a, const CutDelta &dl)
{	// make primal.{data,used,ideal,cnt} and dual.used current again (SURVEY 8(b) coherence rule)
	polytope *P = &a->primal, *D = &a->dual;
	const size_t d = a->dim;
	mirror_reserve(P, P->cnt + dl.n_new);
	if (P->cnt != dl.first_new_slot) die("poly__add_vrtx", "host mirror and device disagree on the slot count");
	// the new slots are contiguous: one block copy of the coordinates, one range of `used` bits
	const size_t s0 = P->cnt, s1 = s0 + dl.n_new;
	if (dl.n_new) memcpy(P->data + s0 * d, dl.coords, (size_t)dl.n_new * d * sizeof(double));
	for (size_t s = s0; s < s1;) {
		const size_t w = s / BTCNT, lo = s % BTCNT, hi = std::min<size_t>(BTCNT, lo + (s1 - s));
		const btstrg mask = (hi == BTCNT ? ~(btstrg)0 : (((btstrg)1 << hi) - 1)) & ~(((btstrg)1 << lo) - 1);
		P->used[w] |= mask;
		s += hi - lo;
	}
	P->cnt = s1;
	for (u32 r = 0; r < dl.n_new; r++) {
		const size_t s = s0 + r;
		if (dl.ideal[r]) ST_BT(P->ideal, s);
		const u32 par = dl.parent_slot[r];
		if (par != B200_NONE && IS_ELEM(P->sltn, par)) {   // copy inherits sltn + pre-image (bslv_poly.c:583-587)
			ST_BT(P->sltn, s);
			memcpy(P->data_primg + s * P->dim_primg, P->data_primg + (size_t)par * P->dim_primg, P->dim_primg * sizeof(double));
		}
	}
	for (u32 i = 0; i < dl.n_dead_entries; i++)
		if (dl.dead_slots[i] != B200_NONE) UNST_BT(P->used, dl.dead_slots[i]);
	for (u32 i = 0; i < dl.n_dead_facets; i++) UNST_BT(D->used, dl.dead_facets[i]);
}

extern "C" int poly__add_vrtx(poly_args *args)
{
	GUARD_BEGIN
	polytope *D = &args->dual;
	const size_t d = args->dim;
	Handle *h = handle_of(D);
	const size_t f = mirror_append(D);                                   // (bslv_poly.c:109-116)
	if (args->ideal) ST_BT(D->ideal, f);
	for (size_t j = 0; j < d; j++) D->data[f * d + j] = args->val[j];
	for (size_t j = 0; j < args->dim_primg_dl; j++) D->data_primg[f * D->dim_primg + j] = args->val_primg_dl[j];
	if (!args->init_data.intlsd) {                                       // queue until initialised (:145)
		queue_push(&args->init_data.queue, f);
		return EXIT_SUCCESS;
	}
	double hp[B200_MAXD + 1];
	((void (*)(double *, int, double *))args->dualV2primalH)(args->val, (int)args->ideal, hp);   // (:119)
	if (h->host_may_have_edited) {
		h->engine->reupload_coords(args->primal.data, args->primal.cnt);
		h->host_may_have_edited = false;
	}
	CutParams P;
	make_params(hp, d, (u32)f, P);
	P.batch_first = (u32)args->primal.cnt;
	CutDelta dl;
	static const bool prof = getenv("B200_PHASES") != nullptr;
	const auto tq0 = std::chrono::steady_clock::now();
	const std::function<void(const CutDelta &)> early = [&](const CutDelta &e) { apply_delta(args, e); };
	h->engine->cut(P, dl, &early);
	const auto tq1 = std::chrono::steady_clock::now();
	if (dl.redundant) {                                                  // (:132-136)
		args->idx = args->primal.cnt;
		UNST_BT(D->used, f);
		return EXIT_FAILURE;
	}
	args->idx = dl.trigger_slot;
	if (!dl.applied_early) apply_delta(args, dl);
	h->lists_current = false;
	if (prof) {
		const auto tq2 = std::chrono::steady_clock::now();
		h->prof_apply_us += std::chrono::duration<double, std::micro>(tq2 - tq1).count();
		h->prof_total_us += std::chrono::duration<double, std::micro>(tq2 - tq0).count();
		if (++h->prof_calls % 2000 == 0) fprintf(stderr, "[b200] poly__add_vrtx: apply_delta %.1f us, cut+apply %.1f us (mean of %llu calls)\n", h->prof_apply_us / h->prof_calls, h->prof_total_us / h->prof_calls, (unsigned long long)h->prof_calls);
	}
	return EXIT_SUCCESS;
	GUARD_END("poly__add_vrtx")
}

extern "C" int poly__get_vrtx(poly_args *args)
{	// first slot that is live and not yet a solution (bslv_poly.c:210-226).  used only goes 1->0 and
	// sltn only 0->1 while slots are append-only, so a monotone cursor replaces the O(S) rescan.
	const polytope *p = &args->primal;
	Handle *h = handle_of(p);
	size_t s = h->get_cursor;
	const size_t n = p->cnt;
	while (s < n) {
		const size_t w = s / BTCNT;
		btstrg m = (p->used[w] & ~p->sltn[w]) >> (s % BTCNT);
		if (m) { s += (size_t)__builtin_ctzll(m); break; }
		s = (w + 1) * BTCNT;
	}
	if (s >= n) { h->get_cursor = n; args->idx = n; return EXIT_FAILURE; }
	h->get_cursor = s;
	args->idx = s;
	args->ideal = (unsigned)IS_ELEM(p->ideal, s);
	for (size_t k = 0; k < p->dim; k++) args->val[k] = p->data[s * p->dim + k];
	return EXIT_SUCCESS;
}

// ------------------------------------------------------------------ start simplex (host, O(d^3))
static double norm2(const double *x, size_t n)
{
	double s = 0;
	for (size_t l = 0; l < n; l++) s += x[l] * x[l];
	return sqrt(s);
}

// bslv__normalise (bslv_poly.c:1030-1060): one modified Gram-Schmidt step.  Row k of Q receives the
// unit residual of x against rows 0..k-1, row k of the packed lower-triangular R the coefficients
// Q_j.x; returns |residual| / |x|, or 0 when the residual is below 1e-6.
extern "C" double bslv__normalise(double *x, double *Q, double *R, size_t k, size_t n)
{
	const double nrm_in = norm2(x, n);
	double *qk = Q + k * n;
	for (size_t l = 0; l < n; l++) qk[l] = x[l];
	for (size_t j = 0; j < k; j++) {
		double s = 0;
		for (size_t l = 0; l < n; l++) s += Q[j * n + l] * qk[l];
		for (size_t l = 0; l < n; l++) qk[l] -= s * Q[j * n + l];
	}
	const double res = norm2(qk, n);
	if (res < 1.0e-6) return 0;
	for (size_t l = 0; l < n; l++) qk[l] /= res;
	for (size_t j = 0; j <= k; j++) {
		double s = 0;
		for (size_t l = 0; l < n; l++) s += Q[j * n + l] * x[l];
		R[k * (k + 1) / 2 + j] = s;
	}
	return res / nrm_in;
}

// poly__poly_initialise (bslv_poly.c:711-787): polyhedron {y : Q_k.y >= alph_k, k<d} as one vertex
// (slot 0) and d extreme directions (slots 1..d); facet perm[k] holds every slot but k, perm[0] is
// the facet at infinity.  Builds the host mirror and uploads the same state to the device.
static void start_simplex(poly_args *a, const double *Q, const double *R, const double *alph, const size_t *perm)
{
	polytope *P = &a->primal;
	const size_t d = P->dim;
	std::vector<double> z(d, 0.0), T(d * d, 0.0);
	auto RR = [&](size_t k, size_t j) { return R[k * (k + 1) / 2 + j]; };
	for (size_t k = 0; k < d; k++) {              // forward substitution R z = alph, R T_k = e_k
		z[k] = alph[k];
		T[k * d + k] = 1.0;
		for (size_t j = 0; j < k; j++) {
			z[k] -= z[j] * RR(k, j);
			for (size_t l = 0; l < d; l++) T[l * d + k] -= T[l * d + j] * RR(k, j);
		}
		z[k] /= RR(k, k);
		for (size_t l = 0; l < d; l++) T[l * d + k] /= RR(k, k);
	}
	const size_t v = mirror_append(P);
	for (size_t k = 0; k < d; k++) {
		double s = 0;
		for (size_t j = 0; j < d; j++) s += z[j] * Q[j * d + k];
		P->data[v * d + k] = s;
	}
	for (size_t k = 0; k < d; k++) {
		const size_t r = mirror_append(P);
		ST_BT(P->ideal, r);
		for (size_t j = 0; j < d; j++) {
			double s = 0;
			for (size_t l = 0; l < d; l++) s += Q[l * d + j] * T[k * d + l];
			P->data[r * d + j] = s;
		}
	}
	const u32 n = (u32)(d + 1), n_facets = (u32)a->dual.cnt;
	std::vector<std::vector<u32>> inc(n), adj(n);
	std::vector<u32> fcnt(n_facets, 0);
	std::vector<u8> ideal(n, 1);
	ideal[0] = 0;
	for (u32 k = 0; k < n; k++)
		for (u32 j = 0; j < n; j++)
			if (j != k) {
				inc[j].push_back((u32)perm[k]);
				fcnt[perm[k]]++;
				adj[k].push_back(j);
			}
	Handle *h = handle_of(P);
	if (!h->engine) h->engine = new CutEngine((int)d);
	h->engine->set_flags(h->flags);
	h->engine->upload_initial(n, P->data, ideal.data(), inc, adj, n_facets, fcnt);
	h->lists_current = false;
}

extern "C" int poly__intl_apprx(poly_args *a)
{
	GUARD_BEGIN
	const size_t d = a->dim;
	poly_list *Qu = &a->init_data.queue, *G = &a->init_data.gnrtrs;
	if (Qu->cnt < d) return EXIT_FAILURE;                                 // (bslv_poly.c:158-159)
	void (*to_hp)(double *, int, double *) = (void (*)(double *, int, double *))a->dualV2primalH;
	std::vector<double> hp((d + 1) * Qu->cnt);
	for (size_t q = 0; q < Qu->cnt; q++)
		to_hp(a->dual.data + Qu->data[q] * d, (int)IS_ELEM(a->dual.ideal, Qu->data[q]), hp.data() + q * (d + 1));
	std::vector<size_t> perm(d + 1, 0);
	size_t chosen = 0;
	while (chosen < d) {                                                  // greedy pivoting (:167-185)
		double best = 0;
		size_t arg = 0;
		for (size_t q = 0; q < Qu->cnt; q++) {
			const double r = bslv__normalise(hp.data() + q * (d + 1), a->init_data.H, a->init_data.R, chosen, d);
			if (best < r) { best = r; arg = q; }
		}
		if (best < 1.0e-10) return EXIT_FAILURE;
		bslv__normalise(hp.data() + arg * (d + 1), a->init_data.H, a->init_data.R, chosen, d);
		a->init_data.alph[chosen] = hp[arg * (d + 1) + d];
		queue_push(G, Qu->data[arg]);
		perm[++chosen] = Qu->data[arg];
		const size_t last = Qu->cnt - 1;
		for (size_t j = 0; j <= d; j++) hp[arg * (d + 1) + j] = hp[last * (d + 1) + j];
		Qu->data[arg] = Qu->data[last];
		Qu->cnt = last;
	}
	start_simplex(a, a->init_data.H, a->init_data.R, a->init_data.alph, perm.data());
	a->init_data.intlsd = 1;
	// queued halfspaces that were not chosen are retired and re-added as fresh dual slots (:190-197)
	for (size_t q = 0; q < Qu->cnt; q++) UNST_BT(a->dual.used, Qu->data[q]);
	// A caller that enumerates vertices queues ALL its halfspaces before this call (cone_vertenum, bslv_algs.c:331-350) and
	// reads the result only after it returns: the re-adds are a known sequence, so they go through the device-resident
	// path (look-ahead + waves, host mirror rebuilt once) -- same slots, same order, same result as the loop below.
	static const size_t batch_from = [] { const char *e = getenv("B200_INIT_BATCH_MIN"); return (size_t)(e ? atol(e) : 32); }();
	const bool batch = Qu->cnt >= batch_from && !a->dim_primg_dl && to_hp == default_dual_to_halfspace;
	if (batch) {
		std::vector<double> vals(Qu->cnt * d);
		std::vector<unsigned char> idl(Qu->cnt);
		for (size_t q = 0; q < Qu->cnt; q++) {
			const size_t src = Qu->data[q];
			memcpy(vals.data() + q * d, a->dual.data + src * d, d * sizeof(double));
			idl[q] = (unsigned char)IS_ELEM(a->dual.ideal, src);
		}
		if (b200_poly_add_batch(a, vals.data(), idl.data(), Qu->cnt, NULL) < 0) die("poly__intl_apprx", b200_last_error());
		if (Qu->cnt) {                  // what the last poly__add_vrtx of the loop would have left in the caller-visible fields
			memcpy(a->val, vals.data() + (Qu->cnt - 1) * d, d * sizeof(double));
			a->ideal = idl[Qu->cnt - 1];
		}
	} else
	for (size_t q = 0; q < Qu->cnt; q++) {
		const size_t src = Qu->data[q];
		for (size_t j = 0; j < d; j++) a->val[j] = a->dual.data[src * d + j];
		a->ideal = (unsigned)IS_ELEM(a->dual.ideal, src);
		poly__add_vrtx(a);
	}
	free(Qu->data); free(G->data);
	free(a->init_data.H); free(a->init_data.R); free(a->init_data.alph);
	Qu->data = G->data = NULL;
	a->init_data.H = a->init_data.R = a->init_data.alph = NULL;
	return EXIT_SUCCESS;
	GUARD_END("poly__intl_apprx")
}

// ------------------------------------------------------------------ lazy host lists
static void point_lists(poly_list *lists, size_t n_slots, const std::vector<size_t> &off, std::vector<size_t> &slab)
{
	for (size_t s = 0; s < n_slots; s++) {
		lists[s].cnt = off[s + 1] - off[s];
		lists[s].blcks = lists[s].cnt;
		lists[s].data = lists[s].cnt ? slab.data() + off[s] : NULL;
	}
}

extern "C" int b200_poly_materialise(poly_args *a)
{
	GUARD_BEGIN
	Handle *h = handle_of(&a->primal);
	if (!h->engine) { mirror_lists_reserve(&a->primal, h, true); mirror_lists_reserve(&a->dual, h, false); return 0; }
	if (h->lists_current && h->lists_cap_p >= a->primal.blcks * SLOTS_PER_BLOCK && h->lists_cap_d >= a->dual.blcks * SLOTS_PER_BLOCK) return 0;
	HostStructure hs;
	h->engine->download_structure(hs);
	polytope *P = &a->primal, *D = &a->dual;
	mirror_lists_reserve(P, h, true);
	mirror_lists_reserve(D, h, false);
	const size_t S = P->cnt, F = D->cnt;
	std::vector<size_t> ioff(S + 1, 0), aoff(S + 1, 0), foff(F + 1, 0);
	auto live = [&](u32 r) { return (hs.live_words[r >> 5] >> (r & 31)) & 1u; };
	for (u32 r = 0; r < hs.nrows; r++) {
		if (!live(r)) continue;
		const u32 s = hs.row_slot[r];
		ioff[s + 1] = hs.inc_len[r];
		aoff[s + 1] = hs.adj_len[r];
		for (u32 q = 0; q < hs.inc_len[r]; q++) foff[hs.inc_pool[hs.inc_off[r] + q] + 1]++;
	}
	for (size_t s = 0; s < S; s++) { ioff[s + 1] += ioff[s]; aoff[s + 1] += aoff[s]; }
	for (size_t f = 0; f < F; f++) foff[f + 1] += foff[f];
	h->slab_pinc.assign(ioff[S], 0);
	h->slab_padj.assign(aoff[S], 0);
	h->slab_dinc.assign(foff[F], 0);
	std::vector<size_t> fill(foff.begin(), foff.end() - 1);
	for (u32 r = 0; r < hs.nrows; r++) {
		if (!live(r)) continue;
		const u32 s = hs.row_slot[r];
		for (u32 q = 0; q < hs.inc_len[r]; q++) {
			const u32 f = hs.inc_pool[hs.inc_off[r] + q];
			h->slab_pinc[ioff[s] + q] = f;
			h->slab_dinc[fill[f]++] = s;
		}
		for (u32 q = 0; q < hs.adj_len[r]; q++) h->slab_padj[aoff[s] + q] = hs.row_slot[hs.adj_pool[hs.adj_off[r] + q]];
	}
	point_lists(P->incidence, S, ioff, h->slab_pinc);
	point_lists(P->adjacence, S, aoff, h->slab_padj);
	point_lists(D->incidence, F, foff, h->slab_dinc);
	h->lists_current = true;
	return 0;
	GUARD_END("b200_poly_materialise")
}

// ------------------------------------------------------------------ combinatorial adjacency on the host lists
// edge_test (bslv_poly.c:467-512) for the end-of-run all-pairs pass on the DUAL polytope
// (poly__update_adjacence(&dual), bslv_algs.c:398,1144,1569).  Not part of the cut step.
static bool list_has(const poly_list *l, size_t x)
{
	for (size_t i = 0; i < l->cnt; i++)
		if (l->data[i] == x) return true;
	return false;
}
static bool adjacent_by_incidence(const polytope *p, size_t u, size_t w)
{
	const poly_list *iu = p->incidence + u, *iw = p->incidence + w;
	std::vector<size_t> mutual;
	for (size_t i = 0; i < iu->cnt; i++)
		if (list_has(iw, iu->data[i])) mutual.push_back(iu->data[i]);
	if (p->dim == 1) return true;
	if (mutual.size() + 1 < p->dim) return false;
	const poly_list *cand = p->dual->incidence + mutual[0];
	for (size_t c = 0; c < cand->cnt; c++) {
		const size_t x = cand->data[c];
		if (x == u || x == w) continue;
		size_t m = 1;
		while (m < mutual.size() && list_has(p->incidence + x, mutual[m])) m++;
		if (m == mutual.size()) return false;
	}
	return true;
}

extern "C" void poly__update_adjacence(polytope *p)
{
	GUARD_BEGIN
	Handle *h = handle_of(p);
	poly_args *a = h->owner;
	b200_poly_materialise(a);
	h->host_may_have_edited = true;      // end-of-run edits of primal.data surround this call
	std::vector<size_t> live;
	for (size_t s = 0; s < p->cnt; s++)
		if (IS_ELEM(p->used, s)) live.push_back(s);
	std::vector<std::vector<size_t>> nb(p->cnt);
	for (size_t s : live)                 // the reference appends to whatever is there (:1003-1004)
		nb[s].assign(p->adjacence[s].data, p->adjacence[s].data + p->adjacence[s].cnt);
	if (p == &a->dual && h->engine) {
		// K6 on the device: rows = live facets, columns = live vertices.  A used facet without a live
		// vertex cannot occur here (clean facet rule), but one with a stale `used` bit is simply isolated.
		std::vector<u32> rank(p->cnt, B200_NONE);
		for (size_t i = 0; i < live.size(); i++) rank[live[i]] = (u32)i;
		std::vector<u32> pa, pb;
		h->engine->dual_adjacency(rank, (u32)live.size(), pa, pb);
		for (size_t q = 0; q < pa.size(); q++) {
			nb[live[pa[q]]].push_back(live[pb[q]]);
			nb[live[pb[q]]].push_back(live[pa[q]]);
		}
		for (size_t s : live) std::sort(nb[s].begin() + p->adjacence[s].cnt, nb[s].end());
		h->lists_current = false;         // the engine compacted its rows: host lists are rebuilt on next use
		b200_poly_materialise(a);
	} else {
		for (size_t i = 0; i < live.size(); i++)
			for (size_t j = i + 1; j < live.size(); j++)
				if (adjacent_by_incidence(p, live[i], live[j])) {
					nb[live[i]].push_back(live[j]);
					nb[live[j]].push_back(live[i]);
				}
	}
	std::vector<size_t> off(p->cnt + 1, 0);
	for (size_t s = 0; s < p->cnt; s++) off[s + 1] = off[s] + nb[s].size();
	std::vector<size_t> &slab = (p == &a->dual) ? h->slab_dadj : h->slab_padj;
	std::vector<size_t> fresh(off[p->cnt]);
	for (size_t s = 0; s < p->cnt; s++) std::copy(nb[s].begin(), nb[s].end(), fresh.begin() + off[s]);
	slab.swap(fresh);
	point_lists(p->adjacence, p->cnt, off, slab);
	GUARD_END("poly__update_adjacence")
}

// ------------------------------------------------------------------ output (formats of bslv_poly.c:314-414)
extern "C" void poly__initialise_permutation(polytope *poly, permutation *prm)
{
	Handle *h = handle_of(poly);
	b200_poly_materialise(h->owner);
	h->host_may_have_edited = true;
	prm->cnt = 0;
	prm->data = (size_t *)malloc(std::max<size_t>(poly->cnt, 1) * sizeof(size_t));
	prm->inv = (size_t *)malloc(std::max<size_t>(poly->cnt, 1) * sizeof(size_t));
	for (size_t s = 0; s < poly->cnt; s++)
		if (IS_ELEM(poly->used, s)) {
			prm->data[prm->cnt] = s;
			prm->inv[s] = prm->cnt++;
		}
}

extern "C" void poly__kill_permutation(permutation *prm)
{
	free(prm->data);
	free(prm->inv);
}

static void end_row(FILE *stream)
{	// the reference overwrites the trailing blank with the newline (bslv_poly.c:352-353)
	fseek(stream, -(long)sizeof(char), SEEK_CUR);
	fprintf(stream, "\n");
}

extern "C" void poly__vrtx2file(polytope *poly, permutation *prm, const char *fname, const char *frmt)
{
	FILE *stream = fname ? fopen(fname, "w") : stdout;
	for (size_t i = 0; i < prm->cnt; i++) {
		const size_t s = prm->data[i];
		fprintf(stream, "%-1.1d ", 1 - (int)IS_ELEM(poly->ideal, s));
		for (size_t j = 0; j < poly->dim; j++) fprintf(stream, frmt ? frmt : "%g ", poly->data[s * poly->dim + j]);
		end_row(stream);
	}
	if (fname) fclose(stream);
}

extern "C" void poly__primg2file(polytope *poly, permutation *prm, const char *fname, const char *frmt)
{
	FILE *stream = fname ? fopen(fname, "w") : stdout;
	for (size_t i = 0; i < prm->cnt; i++) {
		const size_t s = prm->data[i];
		if (!IS_ELEM(poly->sltn, s)) continue;
		for (size_t j = 0; j < poly->dim_primg; j++) fprintf(stream, frmt ? frmt : "%g ", poly->data_primg[s * poly->dim_primg + j]);
		end_row(stream);
	}
	if (fname) fclose(stream);
}

extern "C" void poly__adj2file(polytope *poly, permutation *prm, const char *fname, const char *frmt)
{
	FILE *stream = fname ? fopen(fname, "w") : stdout;
	for (size_t i = 0; i < prm->cnt; i++) {
		const poly_list *l = poly->adjacence + prm->data[i];
		for (size_t q = 0; q < l->cnt; q++) fprintf(stream, frmt ? frmt : "%u ", (unsigned int)prm->inv[l->data[q]]);
		end_row(stream);
	}
	if (fname) fclose(stream);
}

extern "C" void poly__inc2file(polytope *poly, permutation *prm, permutation *prm_dual, const char *fname, const char *frmt)
{
	FILE *stream = fname ? fopen(fname, "w") : stdout;
	for (size_t i = 0; i < prm_dual->cnt; i++) {
		const poly_list *l = poly->dual->incidence + prm_dual->data[i];
		for (size_t q = 0; q < l->cnt; q++) fprintf(stream, frmt ? frmt : "%u ", (unsigned int)prm->inv[l->data[q]]);
		end_row(stream);
	}
	if (fname) fclose(stream);
}

// ------------------------------------------------------------------ swap / plot / check
extern "C" void poly__swap(poly_args *in, poly_args *out)
{	// feed the V-representation of `in` as halfspaces into `out` (bslv_poly.c:836-866)
	b200_poly_materialise(in);
	const size_t d = in->dim;
	for (in->idx = 0; in->idx < in->dual.cnt; in->idx++)
		if (IS_ELEM(in->dual.used, in->idx) && !IS_ELEM(in->dual.ideal, in->idx)) {
			const poly_list *l = in->dual.incidence + in->idx;
			for (size_t q = 0; q < l->cnt; q++) {
				const size_t v = l->data[q];
				for (size_t j = 0; j < d; j++) out->val[j] = in->primal.data[v * d + j];
				out->ideal = (unsigned)IS_ELEM(in->primal.ideal, v);
				poly__add_vrtx(out);
			}
			break;
		}
	poly__intl_apprx(out);
	for (in->idx = 0; in->idx < in->primal.cnt; in->idx++)
		if (IS_ELEM(in->primal.used, in->idx)) {
			for (size_t j = 0; j < d; j++) out->val[j] = in->primal.data[in->idx * d + j];
			out->ideal = (unsigned)IS_ELEM(in->primal.ideal, in->idx);
			poly__add_vrtx(out);
		}
}

extern "C" void poly__plot(polytope *poly, const char *fname)
{	// OFF writer: vertices, then each facet as a cycle of adjacent vertices (bslv_poly.c:868-938)
	permutation perm, dual_perm;
	poly__initialise_permutation(poly, &perm);
	poly__initialise_permutation(poly->dual, &dual_perm);
	poly__update_adjacence(poly->dual);
	FILE *strm = fopen(fname, "w");
	if (!strm) {
		fprintf(stderr, "Error: cannot open %s\n", fname);
	} else {
		fprintf(strm, "OFF\n%zu %zu 0\n\n", perm.cnt, dual_perm.cnt);
		fprintf(strm, "#vertices:\n");
		for (size_t i = 0; i < perm.cnt; i++) {
			for (size_t j = 0; j < poly->dim; j++) fprintf(strm, "%g ", poly->data[perm.data[i] * poly->dim + j]);
			end_row(strm);
		}
		fprintf(strm, "\n#facets:\n");
		for (size_t i = 0; i < dual_perm.cnt; i++) {
			const poly_list *l = poly->dual->incidence + dual_perm.data[i];
			std::vector<size_t> rest(l->data, l->data + l->cnt);
			fprintf(strm, "%zu\t", rest.size());
			bool fault = false;
			while (!rest.empty()) {
				fprintf(strm, "%zu ", perm.inv[rest[0]]);
				if (rest.size() == 1) break;
				size_t nxt = 0;
				for (size_t q = 1; q < rest.size() && !nxt; q++)
					if (list_has(poly->adjacence + rest[0], rest[q])) nxt = q;
				if (!nxt) { fault = true; break; }
				rest[0] = rest[nxt];
				rest[nxt] = rest.back();
				rest.pop_back();
			}
			if (fault) { fprintf(stderr, "Error: Fault in plot.. exiting\n"); break; }
			end_row(strm);
		}
		fclose(strm);
	}
	poly__kill_permutation(&perm);
	poly__kill_permutation(&dual_perm);
}

extern "C" void poly__polyck(poly_args *a)
{	// integrity check of bslv_poly.c:940-990, same messages on stderr
	b200_poly_materialise(a);
	const size_t d = a->dim;
	double hp[B200_MAXD + 1];
	const double eps = 1.0e-6;
	void (*to_hp)(double *, int, double *) = (void (*)(double *, int, double *))a->dualV2primalH;
	for (size_t f = 0; f < a->dual.cnt; f++) {
		if (!IS_ELEM(a->dual.used, f)) continue;
		to_hp(a->dual.data + f * d, (int)IS_ELEM(a->dual.ideal, f), hp);
		const poly_list *l = a->dual.incidence + f;
		for (size_t q = 0; q < l->cnt; q++) {
			const size_t v = l->data[q];
			double s = 0;
			for (size_t j = 0; j < d; j++) s += hp[j] * a->primal.data[v * d + j];
			const double alph = IS_ELEM(a->primal.ideal, v) ? 0 : hp[d];
			if (fabs(s - alph) > eps) fprintf(stderr, "Error:\tHyperplane %zu does not contain vertex %zu.\n", f, v);
			if (!list_has(a->primal.incidence + v, f)) fprintf(stderr, "Error:\tHyperplane %zu, Vertex %zu.\n", f, v);
		}
	}
	for (size_t v = 0; v < a->primal.cnt; v++) {
		if (!IS_ELEM(a->primal.used, v)) continue;
		const poly_list *l = a->primal.adjacence + v;
		for (size_t q = 0; q < l->cnt; q++)
			if (!list_has(a->primal.adjacence + l->data[q], v))
				fprintf(stderr, "Error:\tVertex %zu appears in vertex' %zu adjacence-list, but not vice versa.\n", l->data[q], v);
	}
	for (size_t v = 0; v < a->primal.cnt; v++) {
		if (!IS_ELEM(a->primal.used, v)) continue;
		for (size_t k = 0; k < v; k++)
			if (IS_ELEM(a->primal.used, k) && adjacent_by_incidence(&a->primal, v, k) && !list_has(a->primal.adjacence + v, k))
				fprintf(stderr, "Error:\tVertices %zu and %zu are adjacent (due to their incidence relation), but vertex %zu does not apper in vertex' %zu adjacence-list.\n", v, k, k, v);
	}
}

// ------------------------------------------------------------------ extensions
extern "C" int b200_poly_get_stats(poly_args *a, b200_stats *out)
{
	Handle *h = handle_of(&a->primal);
	memset(out, 0, sizeof *out);
	if (!h->engine) return 1;
	const EngineStats &s = h->engine->stats();
	out->cuts = s.cuts; out->redundant = s.redundant; out->vertex_evals = s.vertex_evals; out->rows_scanned = s.rows_scanned;
	out->minus = s.minus; out->zero = s.zero; out->zero_plus_projected = s.zero_plus_projected;
	out->edge_vertices = s.edge_vertices; out->copies = s.copies; out->pair_tests = s.pair_tests;
	out->new_adjacent_pairs = s.new_adjacent_pairs; out->algorithmic_bytes = s.algorithmic_bytes;
	out->kernel_launches = s.kernel_launches; out->compactions = s.compactions;
	out->live_vertices = h->engine->live_vertices(); out->slots = h->engine->slots(); out->facets = a->dual.cnt;
	out->classify_ms = s.classify_ms; out->cut_ms = s.cut_ms;
	out->waves = s.waves; out->wave_cuts = s.wave_cuts; out->lookahead_passes = s.la_passes; out->sharded_passes = s.sharded_passes;
	out->sharded_cuts = s.sharded_cuts;
	out->sharded_pair_tests = s.sharded_pair_tests;
	return 0;
}

extern "C" int b200_poly_set_flags(poly_args *a, unsigned flags)
{
	Handle *h = handle_of(&a->primal);
	h->flags = flags;
	if (h->engine) h->engine->set_flags(flags);
	return 0;
}

// Rebuild the host mirror in bulk after a device-resident batch (the per-cut deltas were never
// transferred): used/ideal bits and coordinates of every slot created since `first_slot`, dead
// bits of older slots, sltn inheritance through the device-side root journal, dual.used.
static void rebuild_mirror(poly_args *a, Handle *h, size_t first_slot)
{
	polytope *P = &a->primal, *D = &a->dual;
	const size_t d = a->dim;
	MirrorDump m;
	const auto tm0 = std::chrono::steady_clock::now();
	h->engine->download_mirror(m, (u32)D->cnt);
	const auto tm1 = std::chrono::steady_clock::now();
	mirror_reserve(P, m.slot_cnt);
	for (size_t s = P->cnt; s < m.slot_cnt; s++) { UNST_BT(P->used, s); UNST_BT(P->ideal, s); UNST_BT(P->sltn, s); }
	P->cnt = m.slot_cnt;
	for (size_t w = 0; w < (P->cnt + BTCNT - 1) / BTCNT; w++) P->used[w] = 0;
	// Rows are in creation order, so their slots ascend: the row range is cut where the slot crosses a multiple of
	// 64 and each worker owns whole words of the bitsets.  The scatter of 10^6 rows into the slot-indexed mirror is
	// a chain of cache misses for one thread.
	const u32 n = m.nrows;
	auto work = [&](u32 lo, u32 hi) {
		for (u32 r = lo; r < hi; r++) {
			if (!((m.live_words[r >> 5] >> (r & 31)) & 1u)) continue;
			const size_t s = m.row_slot[r];
			ST_BT(P->used, s);
			if (s < first_slot) continue;
			for (size_t j = 0; j < d; j++) P->data[s * d + j] = m.coords_soa[j * (size_t)n + r];
			if ((m.ideal_words[r >> 5] >> (r & 31)) & 1u) ST_BT(P->ideal, s);
		}
	};
	unsigned nt = std::thread::hardware_concurrency();
	nt = nt ? std::min(nt, (unsigned)std::max(1, atoi(getenv("B200_MIRROR_THREADS") ? getenv("B200_MIRROR_THREADS") : "16"))) : 1u;
	if (b200_comm_size() > 1) nt = std::max(std::min(nt, 4u), nt / (unsigned)b200_comm_size());   // one process per GPU: the ranks share the host's cores
	if (n < 65536) nt = 1;
	std::vector<u32> cutp(nt + 1, n);
	cutp[0] = 0;
	for (unsigned t = 1; t < nt; t++) {
		u32 r = (u32)((u64)n * t / nt);
		r = std::max(r, cutp[t - 1]);
		// advance to the first row whose slot starts a new 64-slot word relative to its predecessor
		while (r < n && r > 0 && (m.row_slot[r] / BTCNT) == (m.row_slot[r - 1] / BTCNT)) r++;
		cutp[t] = r;
	}
	if (nt == 1) work(0, n);
	else {
		std::vector<std::thread> th;
		for (unsigned t = 0; t < nt; t++)
			if (cutp[t] < cutp[t + 1]) th.emplace_back(work, cutp[t], cutp[t + 1]);
		for (auto &x : th) x.join();
	}
	const auto tm2 = std::chrono::steady_clock::now();
	if (getenv("B200_PHASES"))
		fprintf(stderr, "[b200] mirror rebuild: download %.1f ms, scatter %.1f ms (%u threads, %u rows)\n", std::chrono::duration<double, std::milli>(tm1 - tm0).count(),
		        std::chrono::duration<double, std::milli>(tm2 - tm1).count(), nt, n);
	// pre-image inheritance through the device-side root journal (rows copied from an on-plane vertex): rare, serial
	for (u32 r = 0; r < n; r++) {
		const u32 root = m.root[r];
		if (root == B200_NONE || !((m.live_words[r >> 5] >> (r & 31)) & 1u)) continue;
		const size_t s = m.row_slot[r];
		if (s < first_slot || !IS_ELEM(P->sltn, root)) continue;
		ST_BT(P->sltn, s);
		memcpy(P->data_primg + s * P->dim_primg, P->data_primg + (size_t)root * P->dim_primg, P->dim_primg * sizeof(double));
	}
	for (size_t f = 0; f < D->cnt; f++) {
		if (m.facet_alive[f]) ST_BT(D->used, f);
		else UNST_BT(D->used, f);
	}
	h->lists_current = false;
}

extern "C" long b200_poly_add_batch_device(poly_args *a, const double *d_vals, const unsigned char *d_ideal, size_t n, int *rc_out)
{
	GUARD_BEGIN
	Handle *h = handle_of(&a->primal);
	if (!a->init_data.intlsd || !h->engine) { b200_set_error("b200_poly_add_batch_device: call poly__intl_apprx first"); return -1; }
	if ((void (*)(double *, int, double *))a->dualV2primalH != default_dual_to_halfspace) {
		b200_set_error("b200_poly_add_batch_device: only the default dual->halfspace callback (cone_polar) is evaluated on the device");
		return -1;
	}
	if (a->dim_primg_dl) { b200_set_error("b200_poly_add_batch_device: dual pre-images are not supported in batch mode"); return -1; }
	polytope *D = &a->dual;
	const size_t d = a->dim, first_slot = a->primal.cnt, f0 = D->cnt;
	if (h->host_may_have_edited) {
		h->engine->reupload_coords(a->primal.data, a->primal.cnt);
		h->host_may_have_edited = false;
	}
	const auto t0 = std::chrono::steady_clock::now();
	for (size_t i = 0; i < n; i++) mirror_append(D);          // halfspace i becomes dual slot f0 + i whatever its fate (bslv_poly.c:109-116)
	const long cuts = h->engine->cut_batch_from_device(d_vals, d_ideal, n, (u32)f0, (u32)first_slot, rc_out);
	const auto t1 = std::chrono::steady_clock::now();
	// dual rows: the points themselves (host copy of the device inputs) and their ideal flags
	std::vector<double> hv(n * d);
	std::vector<unsigned char> hi(n, 0);
	h->engine->device_download(hv.data(), d_vals, n * d * sizeof(double));
	if (d_ideal) h->engine->device_download(hi.data(), d_ideal, n);
	for (size_t i = 0; i < n; i++) {
		memcpy(D->data + (f0 + i) * d, hv.data() + i * d, d * sizeof(double));
		if (hi[i]) ST_BT(D->ideal, f0 + i);
	}
	rebuild_mirror(a, h, first_slot);
	a->idx = a->primal.cnt;
	if (getenv("B200_PHASES"))
		fprintf(stderr, "[b200] batch of %zu: cuts %.1f ms, mirror rebuild %.1f ms\n", n, std::chrono::duration<double, std::milli>(t1 - t0).count(),
		        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
	return cuts;
	GUARD_END("b200_poly_add_batch_device")
}

// the loop a C caller (bslv_algs.c) writes: one poly__add_vrtx per halfspace, mirror coherent after each
extern "C" long b200_poly_add_each(poly_args *a, const double *vals, const unsigned char *ideal, size_t n, int *rc_out)
{
	long cuts = 0;
	for (size_t i = 0; i < n; i++) {
		for (size_t j = 0; j < a->dim; j++) a->val[j] = vals[i * a->dim + j];
		a->ideal = ideal ? ideal[i] : 0;
		const int rc = poly__add_vrtx(a);
		if (rc_out) rc_out[i] = rc;
		cuts += (rc == EXIT_SUCCESS);
	}
	return cuts;
}

extern "C" long b200_poly_add_batch(poly_args *a, const double *vals, const unsigned char *ideal, size_t n, int *rc_out)
{
	GUARD_BEGIN
	Handle *h = handle_of(&a->primal);
	const bool device_path = a->init_data.intlsd && h->engine && !a->dim_primg_dl &&
	                         (void (*)(double *, int, double *))a->dualV2primalH == default_dual_to_halfspace;
	if (!device_path) return b200_poly_add_each(a, vals, ideal, n, rc_out);   // generic callback / not yet initialised
	double *dv = (double *)h->engine->device_alloc(n * a->dim * sizeof(double));
	unsigned char *di = ideal ? (unsigned char *)h->engine->device_alloc(n) : nullptr;
	h->engine->device_upload(dv, vals, n * a->dim * sizeof(double));
	if (ideal) h->engine->device_upload(di, ideal, n);
	const long cuts = b200_poly_add_batch_device(a, dv, di, n, rc_out);
	h->engine->device_free(dv);
	h->engine->device_free(di);
	return cuts;
	GUARD_END("b200_poly_add_batch")
}

extern "C" int b200_poly_reserve(poly_args *a, size_t vertices, size_t incidence_entries, size_t adjacency_entries)
{
	GUARD_BEGIN
	Handle *h = handle_of(&a->primal);
	if (!h->engine) h->engine = new CutEngine((int)a->dim);
	h->engine->reserve(vertices, incidence_entries, adjacency_entries);
	mirror_reserve(&a->primal, vertices);
	return 0;
	GUARD_END("b200_poly_reserve")
}

extern "C" double b200_poly_classify_bench(poly_args *a, const double *hp, int iters, int flush_l2)
{
	GUARD_BEGIN
	Handle *h = handle_of(&a->primal);
	if (!h->engine) return -1.0;
	CutParams P;
	make_params(hp, a->dim, (u32)a->dual.cnt, P);     // an id one past the last facet: scratch only
	P.batch_first = (u32)a->primal.cnt;
	h->engine->compact();            // measure on dense rows, as right after a compaction
	return h->engine->classify_bench(P, iters, flush_l2);
	GUARD_END("b200_poly_classify_bench")
}

// multi-GPU set-up (one process per GPU): rank 0 makes an id, the launcher broadcasts it, every rank
// calls b200_comm_init BEFORE poly__initialise.  All ranks then issue the same call sequence.
extern "C" int b200_comm_unique_id(char out[128]) { return b200_comm_make_id(out); }
extern "C" int b200_comm_init(int rank, int nranks, const char id[128]) { return b200_comm_start(rank, nranks, id); }
extern "C" void b200_comm_finalize(void) { b200_comm_stop(); }
extern "C" int b200_comm_set_allgather_callback(void (*fn)(const void *, void *, size_t)) { return b200_comm_set_callback(fn); }

extern "C" int b200_set_device(int device) { return b200_select_device(device); }
extern "C" int b200_device_count(void) { return b200_num_devices(); }
extern "C" const char *b200_version(void) { return "bensolve_b200 0.1 (sm_100a)"; }
extern "C" const char *b200_last_error(void) { return b200_get_error(); }
