/*
 * Stand-in for <glpk.h> (GLPK is not installed in this image and there is no network).
 *
 * Declares exactly the 28 functions and the constants bensolve's unchanged bslv_lp.c uses
 * (bslv_lp.c:21-325), with GLPK's real constant values (bslv_lp.c:34-36 indexes a table by them).
 * The implementation (glpk_shim.c) only stores the problem; glp_simplex hands it to a solver
 * callback registered by the host process (tools/run_bensolve.py uses scipy's HiGHS).
 * Test / integration infrastructure: lets the reference CLI run here so that real Benson cut
 * sequences can be recorded and the B200 engine can be exercised end to end.
 */
#ifndef GLPK_SHIM_H
#define GLPK_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif

#define GLP_FR 1
#define GLP_LO 2
#define GLP_UP 3
#define GLP_DB 4
#define GLP_FX 5
#define GLP_PRIMAL 1
#define GLP_DUALP 2
#define GLP_DUAL 3
#define GLP_UNDEF 1
#define GLP_FEAS 2
#define GLP_INFEAS 3
#define GLP_NOFEAS 4
#define GLP_OPT 5
#define GLP_UNBND 6
#define GLP_MSG_OFF 0
#define GLP_MSG_ERR 1
#define GLP_MSG_ON 2
#define GLP_MSG_ALL 3
#define GLP_ON 1
#define GLP_OFF 0

typedef struct glp_prob glp_prob;
typedef struct { int msg_lev, meth, pricing, r_test; double tol_bnd, tol_dj, tol_piv, obj_ll, obj_ul; int it_lim, tm_lim, out_frq, out_dly, presolve; double foo_bar[36]; } glp_smcp;

void glp_init_smcp(glp_smcp *parm);
glp_prob *glp_create_prob(void);
int glp_add_rows(glp_prob *P, int nrs);
int glp_add_cols(glp_prob *P, int ncs);
void glp_load_matrix(glp_prob *P, int ne, const int ia[], const int ja[], const double ar[]);
int glp_get_num_rows(glp_prob *P);
int glp_get_num_cols(glp_prob *P);
void glp_del_rows(glp_prob *P, int nrs, const int num[]);
void glp_del_cols(glp_prob *P, int ncs, const int num[]);
void glp_std_basis(glp_prob *P);
void glp_copy_prob(glp_prob *dest, glp_prob *prob, int names);
void glp_set_row_bnds(glp_prob *P, int i, int type, double lb, double ub);
void glp_set_col_bnds(glp_prob *P, int j, int type, double lb, double ub);
void glp_set_mat_row(glp_prob *P, int i, int len, const int ind[], const double val[]);
void glp_set_obj_coef(glp_prob *P, int j, double coef);
int glp_simplex(glp_prob *P, const glp_smcp *parm);
int glp_get_status(glp_prob *P);
int glp_get_prim_stat(glp_prob *P);
int glp_get_dual_stat(glp_prob *P);
double glp_get_row_prim(glp_prob *P, int i);
double glp_get_col_prim(glp_prob *P, int j);
double glp_get_row_dual(glp_prob *P, int i);
double glp_get_col_dual(glp_prob *P, int j);
double glp_get_obj_val(glp_prob *P);
void glp_delete_prob(glp_prob *P);
int glp_free_env(void);
int glp_write_prob(glp_prob *P, int flags, const char *fname);
int glp_write_sol(glp_prob *P, const char *fname);

/* shim-only: the host process registers the solver */
typedef int (*glp_shim_solver)(glp_prob *P, int meth);
void glp_shim_set_solver(glp_shim_solver fn);

#ifdef __cplusplus
}
#endif
#endif
