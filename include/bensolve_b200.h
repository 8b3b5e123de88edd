/*
 * bensolve_b200 -- C ABI of the B200 polyhedral cut engine.
 *
 * This header declares exactly what bensolve's unchanged host code (bslv_algs.c) binds for the
 * polyhedron engine: the struct layouts of bslv_poly.h:49-88 and the poly__* entry points of
 * bslv_poly.h:90-118.  libbslv_poly_b200.so exports them with the reference's names, argument
 * meaning and return codes, so it links in place of bslv_poly.o.  The cut itself
 * (poly__add_vrtx -> poly__cut -> edge_test in the reference, bslv_poly.c:104-151, 562-709,
 * 467-512) runs as hand-written sm_100a kernels; there is no CPU fallback for it.
 *
 * Every function is extern "C", takes plain pointers and sizes and owns no torch types.
 * b200_* functions are extensions (batch / device-resident paths, statistics, multi-GPU set-up);
 * the reference has no counterpart for them.
 */
#ifndef BENSOLVE_B200_H
#define BENSOLVE_B200_H

#include <limits.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- bitset helpers, identical to bslv_poly.h:40-45 (the caller uses them on our arrays) ---- */
typedef size_t btstrg;
typedef btstrg vrtx_strg;
#ifndef BTCNT
#define BTCNT (CHAR_BIT * sizeof(btstrg))
#define ST_BT(lst, idx) (*((lst) + (idx) / BTCNT) |= (btstrg)1 << (idx) % BTCNT)
#define UNST_BT(lst, idx) (*((lst) + (idx) / BTCNT) &= ~((btstrg)1 << (idx) % BTCNT))
#define IS_ELEM(lst, idx) ((btstrg)1U & (*((lst) + (idx) / BTCNT) >> (idx) % BTCNT))
#endif
#ifndef POLY_EPS
#define POLY_EPS 1e-9 /* bslv_poly.h:47 */
#endif

/* ---- layouts: bslv_poly.h:49-53 (24 B), :55-69 (112 B), :71-82 (392 B), :84-88 ---- */
typedef struct poly_list_strct {
	size_t cnt;
	size_t blcks;
	size_t *data;
} poly_list;

typedef struct polytope_strct {
	size_t dim, dim_primg;
	size_t cnt;   /* slots ever created; dead slots are never reused */
	size_t blcks; /* allocated capacity in units of 64 slots */
	double *ip;   /* dead field in the reference (:233, :725); we keep our engine handle here */
	double *data; /* AoS [cnt][dim], host-coherent after every poly__add_vrtx / poly__intl_apprx */
	double *data_primg; /* [cnt][dim_primg], host-authoritative */
	poly_list *adjacence; /* materialised lazily (writers, polyck, update_adjacence, swap, plot) */
	poly_list *incidence; /* idem */
	vrtx_strg *ideal;
	vrtx_strg *used;
	vrtx_strg *sltn; /* host-authoritative: the caller sets bits directly (bslv_algs.c:1076) */
	struct polytope_strct *dual;
	void (*v2h)(double *, int, double *);
} polytope;

typedef struct {
	size_t dim, dim_primg_prml, dim_primg_dl;
	unsigned int ideal : 1;
	size_t idx;
	double *val, *val_primg_prml, *val_primg_dl;
	double eps; /* written, never read (bslv_poly.c:45) */
	polytope primal;
	polytope dual;
	void (*primalV2dualH)();
	void (*dualV2primalH)(); /* (double *dual_vertex, int is_dir, double *hp_out[dim+1]) */
	struct {
		double *H, *R, *alph;
		poly_list queue, gnrtrs;
		unsigned int intlsd : 1;
	} init_data;
} poly_args;

typedef struct {
	size_t cnt;
	size_t *data;
	size_t *inv;
} permutation;

/* ---- the boundary: the 16 symbols bslv_algs.o imports ---- */
void poly__set_default_args(poly_args *args, size_t dim);   /* replaces bslv_poly.c:41-53   */
void poly__initialise(poly_args *);                         /* replaces bslv_poly.c:55-102  */
int poly__add_vrtx(poly_args *);                            /* replaces bslv_poly.c:104-151 (+ :562-709, :467-512): THE CUT */
int poly__intl_apprx(poly_args *);                          /* replaces bslv_poly.c:153-208 (+ :711-787, :1030-1060) */
int poly__get_vrtx(poly_args *);                            /* replaces bslv_poly.c:210-226 */
void poly__kill(poly_args *);                               /* replaces bslv_poly.c:258-294 */
void poly__update_adjacence(polytope *);                    /* replaces bslv_poly.c:992-1010 */
void poly__swap(poly_args *, poly_args *);                  /* replaces bslv_poly.c:836-866 */
void poly__plot(polytope *, const char *);                  /* replaces bslv_poly.c:868-938 */
void poly__polyck(poly_args *poly);                         /* replaces bslv_poly.c:940-990 */
void poly__initialise_permutation(polytope *, permutation *);                               /* :314-330 */
void poly__kill_permutation(permutation *);                                                 /* :332-339 */
void poly__vrtx2file(polytope *, permutation *, const char *, const char *);                /* :341-360 */
void poly__primg2file(polytope *, permutation *, const char *, const char *);               /* :362-380 */
void poly__adj2file(polytope *, permutation *, const char *, const char *);                 /* :382-397 */
void poly__inc2file(polytope *, permutation *, permutation *, const char *, const char *);  /* :399-414 */

/* ---- extensions (no reference counterpart) ---- */

/* Bring primal/dual .incidence and .adjacence host lists up to date from device state.
 * Called internally by the writers, polyck, update_adjacence, swap and plot.  0 on success. */
int b200_poly_materialise(poly_args *);

/* Batched, device-resident form of a run of poly__add_vrtx calls (default cone_polar callback
 * semantics: halfspace vals[i].y >= -1, or >= 0 where ideal[i]).  vals is HOST memory [n][dim],
 * ideal may be NULL (all zero).  rc_out (may be NULL) receives the n return codes.  The host
 * mirror is made coherent once, when the call returns.  Returns the number of non-redundant cuts,
 * or -1 on error. */
long b200_poly_add_batch(poly_args *, const double *vals, const unsigned char *ideal, size_t n, int *rc_out);

/* The loop a C caller writes around poly__add_vrtx (as bslv_algs.c does): for each i copy vals[i] into args->val,
 * set args->ideal, call poly__add_vrtx; the host mirror is coherent after every call.  Works with any callback.
 * Returns the number of non-redundant cuts. */
long b200_poly_add_each(poly_args *, const double *vals, const unsigned char *ideal, size_t n, int *rc_out);

/* Same with the dual points already resident in HBM (device pointer, row-major [n][dim]). */
long b200_poly_add_batch_device(poly_args *, const double *d_vals, const unsigned char *d_ideal, size_t n, int *rc_out);

/* Pre-size device and host storage (vertices ever created, incidence and adjacency entries) so
 * that no re-allocation happens inside a timed region.  May be called right after poly__initialise. */
int b200_poly_reserve(poly_args *, size_t vertices, size_t incidence_entries, size_t adjacency_entries);

/* Measurement hook: launch K1 (classify) alone `iters` times against halfspace hp[dim+1] without
 * mutating the polytope, flushing L2 before each launch when flush_l2 != 0.  Returns the mean
 * CUDA-event time of one launch in milliseconds (< 0 on error). */
double b200_poly_classify_bench(poly_args *, const double *hp, int iters, int flush_l2);

/* Cumulative statistics of one engine since poly__initialise. */
typedef struct {
	uint64_t cuts;             /* non-redundant poly__add_vrtx calls */
	uint64_t redundant;        /* calls that returned EXIT_FAILURE */
	uint64_t vertex_evals;     /* live vertices classified (K1) */
	uint64_t rows_scanned;     /* device rows K1 walked (live + not yet compacted dead) */
	uint64_t minus, zero, zero_plus_projected;
	uint64_t edge_vertices, copies;
	uint64_t pair_tests, new_adjacent_pairs;
	uint64_t algorithmic_bytes; /* SURVEY 8(d) formula, summed over cuts */
	uint64_t kernel_launches;
	uint64_t compactions;
	uint64_t live_vertices, slots, facets; /* current */
	double classify_ms;        /* sum of CUDA-event times of K1 when timing is enabled */
	double cut_ms;             /* sum of CUDA-event times of whole cuts when timing is enabled */
	/* device-resident batches (wave path): waves run, cuts they carried out, look-ahead passes over the coordinates,
	 * and how many of those passes were split across the ranks (several GPUs) */
	uint64_t waves, wave_cuts, lookahead_passes, sharded_passes;
	uint64_t sharded_cuts;     /* per-call path: cuts whose K1 was split across the ranks */
	uint64_t sharded_pair_tests; /* waves whose pair test was split across the ranks */
} b200_stats;
int b200_poly_get_stats(poly_args *, b200_stats *out);
/* flags: bit0 = time every K1 launch and every cut with CUDA events (adds two syncs per cut);
 *        bit1 = compact device rows as soon as one is dead (test hook);
 *        bit2 = always use the multi-kernel path, never the single-CTA tail (test hook) */
int b200_poly_set_flags(poly_args *, unsigned flags);

/* Multi-GPU (one process per GPU, state replicated, K1 sharded by row range, visited lists merged by
 * an NCCL all-gather per cut).  Rank 0 calls b200_comm_unique_id, the launcher broadcasts the 128
 * bytes, every rank calls b200_comm_init before poly__initialise; afterwards all ranks issue the
 * same poly__* call sequence with the same arguments and hold identical state.
 * b200_comm_set_allgather_callback exists for the host-side test double only (the product returns 1). */
int b200_comm_unique_id(char out[128]);
int b200_comm_init(int rank, int nranks, const char id[128]);
void b200_comm_finalize(void);
int b200_comm_set_allgather_callback(void (*fn)(const void *send, void *recv, size_t bytes_per_rank));

/* Library-wide: device selection (before the first poly__initialise), version, last error text. */
int b200_set_device(int device);
int b200_device_count(void);
const char *b200_version(void);
const char *b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* BENSOLVE_B200_H */
