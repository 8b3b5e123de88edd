/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of bensolve's polyhedron cut path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (bensolve_b200/libbslv_poly_b200.so) never links or calls it.
 *
 * Parity pinning: the reference ships no golden vectors for this path; this
 * restatement is pinned against the UNMODIFIED reference object oracle/_ref/libref_poly.so
 * (built from /root/reference/bslv_poly.c by oracle/Makefile) on the traces of tests/ and against
 * the fixtures in tests/golden/ that were generated from that object (tests/golden/make_golden.py).
 *
 * Struct layouts restate bslv_poly.h:49-88 (LP64: poly_list 24 B, polytope 112 B, poly_args 392 B)
 * so the same ctypes reader (bensolve_b200/capi.py) drives the reference, this file and the CUDA
 * engine.
 */
#ifndef POLY_ORACLE_H
#define POLY_ORACLE_H
#include <stddef.h>

typedef struct { size_t cnt, blcks; size_t *data; } poly_list;            /* bslv_poly.h:49-53 */

typedef struct polytope_s {                                               /* bslv_poly.h:55-69 */
	size_t dim, dim_primg, cnt, blcks;
	double *ip, *data, *data_primg;
	poly_list *adjacence, *incidence;
	size_t *ideal, *used, *sltn;          /* 64-bit bitsets: word idx/64, bit idx%64 (:40-45) */
	struct polytope_s *dual;
	void (*v2h)(double *, int, double *);
} polytope;

typedef struct {                                                          /* bslv_poly.h:71-82 */
	size_t dim, dim_primg_prml, dim_primg_dl;
	unsigned int ideal : 1;
	size_t idx;
	double *val, *val_primg_prml, *val_primg_dl;
	double eps;
	polytope primal, dual;
	void (*primalV2dualH)(double *, int, double *);
	void (*dualV2primalH)(double *, int, double *);
	struct { double *H, *R, *alph; poly_list queue, gnrtrs; unsigned int intlsd : 1; } init_data;
} poly_args;

void poly__set_default_args(poly_args *, size_t dim);                     /* bslv_poly.c:41-53  */
void poly__initialise(poly_args *);                                       /* bslv_poly.c:55-102 */
int poly__add_vrtx(poly_args *);                                          /* bslv_poly.c:104-151 */
int poly__intl_apprx(poly_args *);                                        /* bslv_poly.c:153-208 */
int poly__get_vrtx(poly_args *);                                          /* bslv_poly.c:210-226 */
void poly__update_adjacence(polytope *);                                  /* bslv_poly.c:992-1010 */
void poly__kill(poly_args *);                                             /* bslv_poly.c:258-294 */

/* diagnostics for the tests: statistics of the last cut */
typedef struct {
	size_t n_minus, n_zero, n_zero_plus_projected, n_edge_vertices, n_copies;
	size_t n_nonplus_unreached;   /* non-PLUS live vertices NOT reachable from v0 (convexity check) */
	size_t n_pair_tests, n_new_adjacent_pairs;
} oracle_cut_stats;
const oracle_cut_stats *oracle_last_cut_stats(void);

#endif
