/* Problem container behind tools/lpshim/glpk.h -- see the header. 1-based indices as in GLPK. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "glpk.h"

struct glp_prob {
	int m, n, cap_m, cap_n;
	int *rtype, *ctype;              /* [1..] */
	double *rlb, *rub, *clb, *cub, *obj; /* obj[0] = constant term */
	int *rlen;
	int **rind;
	double **rval;
	int status, pstat, dstat;
	double objval;
	double *rprim, *rdual, *cprim, *cdual;
};

static glp_shim_solver g_solver;
void glp_shim_set_solver(glp_shim_solver fn) { g_solver = fn; }

void glp_init_smcp(glp_smcp *parm) { memset(parm, 0, sizeof *parm); parm->msg_lev = GLP_MSG_ALL; parm->meth = GLP_PRIMAL; }

glp_prob *glp_create_prob(void)
{
	glp_prob *P = (glp_prob *)calloc(1, sizeof *P);
	P->obj = (double *)calloc(1, sizeof(double));
	P->status = P->pstat = P->dstat = GLP_UNDEF;
	return P;
}

#define GROW(ptr, type, n) ptr = (type *)realloc(ptr, (size_t)(n) * sizeof(type))
static void grow_rows(glp_prob *P, int m)
{
	if (m + 1 <= P->cap_m) return;
	int cap = m + 1 + P->cap_m;
	GROW(P->rtype, int, cap); GROW(P->rlb, double, cap); GROW(P->rub, double, cap); GROW(P->rlen, int, cap);
	GROW(P->rind, int *, cap); GROW(P->rval, double *, cap); GROW(P->rprim, double, cap); GROW(P->rdual, double, cap);
	for (int i = P->cap_m; i < cap; i++) { P->rlen[i] = 0; P->rind[i] = NULL; P->rval[i] = NULL; }
	P->cap_m = cap;
}
static void grow_cols(glp_prob *P, int n)
{
	if (n + 1 <= P->cap_n) return;
	int cap = n + 1 + P->cap_n;
	GROW(P->ctype, int, cap); GROW(P->clb, double, cap); GROW(P->cub, double, cap); GROW(P->obj, double, cap);
	GROW(P->cprim, double, cap); GROW(P->cdual, double, cap);
	P->cap_n = cap;
}

int glp_add_rows(glp_prob *P, int nrs)
{
	int first = P->m + 1;
	grow_rows(P, P->m + nrs);
	for (int i = first; i < first + nrs; i++) {      /* GLPK default: free row, empty */
		P->rtype[i] = GLP_FR; P->rlb[i] = P->rub[i] = 0; P->rlen[i] = 0; P->rprim[i] = P->rdual[i] = 0;
		free(P->rind[i]); free(P->rval[i]); P->rind[i] = NULL; P->rval[i] = NULL;
	}
	P->m += nrs;
	return first;
}
int glp_add_cols(glp_prob *P, int ncs)
{
	int first = P->n + 1;
	grow_cols(P, P->n + ncs);
	for (int j = first; j < first + ncs; j++) {      /* GLPK default: fixed at zero, objective 0 */
		P->ctype[j] = GLP_FX; P->clb[j] = P->cub[j] = 0; P->obj[j] = 0; P->cprim[j] = P->cdual[j] = 0;
	}
	P->n += ncs;
	return first;
}
static void set_row(glp_prob *P, int i, int len, const int *ind, const double *val)
{
	free(P->rind[i]); free(P->rval[i]);
	P->rind[i] = (int *)malloc((size_t)(len ? len : 1) * sizeof(int));
	P->rval[i] = (double *)malloc((size_t)(len ? len : 1) * sizeof(double));
	int k = 0;
	for (int t = 1; t <= len; t++)
		if (val[t] != 0.0) { P->rind[i][k] = ind[t]; P->rval[i][k] = val[t]; k++; }
	P->rlen[i] = k;
}
void glp_load_matrix(glp_prob *P, int ne, const int ia[], const int ja[], const double ar[])
{
	int *cnt = (int *)calloc((size_t)P->m + 2, sizeof(int));
	for (int k = 1; k <= ne; k++) cnt[ia[k]]++;
	for (int i = 1; i <= P->m; i++) {
		free(P->rind[i]); free(P->rval[i]);
		P->rind[i] = (int *)malloc((size_t)(cnt[i] ? cnt[i] : 1) * sizeof(int));
		P->rval[i] = (double *)malloc((size_t)(cnt[i] ? cnt[i] : 1) * sizeof(double));
		P->rlen[i] = 0;
	}
	for (int k = 1; k <= ne; k++)
		if (ar[k] != 0.0) { int i = ia[k]; P->rind[i][P->rlen[i]] = ja[k]; P->rval[i][P->rlen[i]] = ar[k]; P->rlen[i]++; }
	free(cnt);
}
void glp_set_mat_row(glp_prob *P, int i, int len, const int ind[], const double val[]) { set_row(P, i, len, ind, val); }
int glp_get_num_rows(glp_prob *P) { return P->m; }
int glp_get_num_cols(glp_prob *P) { return P->n; }

void glp_del_rows(glp_prob *P, int nrs, const int num[])
{
	char *del = (char *)calloc((size_t)P->m + 2, 1);
	for (int k = 1; k <= nrs; k++) del[num[k]] = 1;
	int w = 1;
	for (int i = 1; i <= P->m; i++) {
		if (del[i]) { free(P->rind[i]); free(P->rval[i]); P->rind[i] = NULL; P->rval[i] = NULL; P->rlen[i] = 0; continue; }
		if (w != i) {
			P->rtype[w] = P->rtype[i]; P->rlb[w] = P->rlb[i]; P->rub[w] = P->rub[i];
			P->rlen[w] = P->rlen[i]; P->rind[w] = P->rind[i]; P->rval[w] = P->rval[i];
			P->rind[i] = NULL; P->rval[i] = NULL; P->rlen[i] = 0;
		}
		w++;
	}
	P->m = w - 1;
	free(del);
}
void glp_del_cols(glp_prob *P, int ncs, const int num[])
{
	int *map = (int *)calloc((size_t)P->n + 2, sizeof(int));
	for (int k = 1; k <= ncs; k++) map[num[k]] = -1;
	int w = 1;
	for (int j = 1; j <= P->n; j++) {
		if (map[j] < 0) continue;
		map[j] = w;
		P->ctype[w] = P->ctype[j]; P->clb[w] = P->clb[j]; P->cub[w] = P->cub[j]; P->obj[w] = P->obj[j];
		w++;
	}
	for (int i = 1; i <= P->m; i++) {
		int k = 0;
		for (int t = 0; t < P->rlen[i]; t++)
			if (map[P->rind[i][t]] > 0) { P->rind[i][k] = map[P->rind[i][t]]; P->rval[i][k] = P->rval[i][t]; k++; }
		P->rlen[i] = k;
	}
	P->n = w - 1;
	free(map);
}
void glp_std_basis(glp_prob *P) { (void)P; }
void glp_copy_prob(glp_prob *dest, glp_prob *src, int names)
{
	(void)names;
	for (int i = 1; i <= src->m && i <= dest->m; i++) {
		dest->rtype[i] = src->rtype[i]; dest->rlb[i] = src->rlb[i]; dest->rub[i] = src->rub[i];
		free(dest->rind[i]); free(dest->rval[i]);
		dest->rind[i] = (int *)malloc((size_t)(src->rlen[i] ? src->rlen[i] : 1) * sizeof(int));
		dest->rval[i] = (double *)malloc((size_t)(src->rlen[i] ? src->rlen[i] : 1) * sizeof(double));
		memcpy(dest->rind[i], src->rind[i], (size_t)src->rlen[i] * sizeof(int));
		memcpy(dest->rval[i], src->rval[i], (size_t)src->rlen[i] * sizeof(double));
		dest->rlen[i] = src->rlen[i];
	}
	for (int j = 0; j <= src->n && j <= dest->n; j++) {
		dest->obj[j] = src->obj[j];
		if (j) { dest->ctype[j] = src->ctype[j]; dest->clb[j] = src->clb[j]; dest->cub[j] = src->cub[j]; }
	}
}
void glp_set_row_bnds(glp_prob *P, int i, int type, double lb, double ub) { P->rtype[i] = type; P->rlb[i] = lb; P->rub[i] = ub; }
void glp_set_col_bnds(glp_prob *P, int j, int type, double lb, double ub) { P->ctype[j] = type; P->clb[j] = lb; P->cub[j] = ub; }
void glp_set_obj_coef(glp_prob *P, int j, double coef) { if (j <= P->n) P->obj[j] = coef; }

int glp_simplex(glp_prob *P, const glp_smcp *parm)
{
	if (!g_solver) { fprintf(stderr, "glpk shim: no solver registered\n"); abort(); }
	P->status = P->pstat = P->dstat = GLP_UNDEF;
	return g_solver(P, parm ? parm->meth : GLP_PRIMAL);
}
int glp_get_status(glp_prob *P) { return P->status; }
int glp_get_prim_stat(glp_prob *P) { return P->pstat; }
int glp_get_dual_stat(glp_prob *P) { return P->dstat; }
double glp_get_row_prim(glp_prob *P, int i) { return P->rprim[i]; }
double glp_get_col_prim(glp_prob *P, int j) { return P->cprim[j]; }
double glp_get_row_dual(glp_prob *P, int i) { return P->rdual[i]; }
double glp_get_col_dual(glp_prob *P, int j) { return P->cdual[j]; }
double glp_get_obj_val(glp_prob *P) { return P->objval; }
void glp_delete_prob(glp_prob *P)
{
	if (!P) return;
	for (int i = 0; i < P->cap_m; i++) { free(P->rind[i]); free(P->rval[i]); }
	free(P->rtype); free(P->rlb); free(P->rub); free(P->rlen); free(P->rind); free(P->rval); free(P->rprim); free(P->rdual);
	free(P->ctype); free(P->clb); free(P->cub); free(P->obj); free(P->cprim); free(P->cdual);
	free(P);
}
int glp_free_env(void) { return 0; }
int glp_write_prob(glp_prob *P, int flags, const char *fname) { (void)P; (void)flags; (void)fname; return 0; }
int glp_write_sol(glp_prob *P, const char *fname) { (void)P; (void)fname; return 0; }
