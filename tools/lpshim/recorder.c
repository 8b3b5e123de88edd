/* Trace recorder (build container only): interposes the poly__* entry points bslv_algs.o calls,
 * logs every call at the `val` level together with the halfspace the
 * caller's callback derives from it, and forwards to the next definition in the link map (the
 * unmodified reference engine).  One JSON object per line on the file named by $BSLV_TRACE. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bensolve_b200.h"

static FILE *out(void)
{
	static FILE *f;
	if (!f) {
		const char *p = getenv("BSLV_TRACE");
		f = p ? fopen(p, "w") : NULL;
		if (!f) { fprintf(stderr, "recorder: set BSLV_TRACE\n"); abort(); }
	}
	return f;
}
static void hexvec(FILE *f, const char *key, const double *v, size_t n)
{
	fprintf(f, "\"%s\":[", key);
	for (size_t i = 0; i < n; i++) fprintf(f, "%s\"%a\"", i ? "," : "", v[i]);
	fprintf(f, "]");
}
typedef void (*void_fn)(poly_args *);
typedef int (*int_fn)(poly_args *);
static void *engine(void)
{	/* RTLD_NEXT does not see libraries the host process dlopen()ed; open the engine by path */
	static void *h;
	if (!h) {
		const char *p = getenv("BSLV_ENGINE_SO");
		h = p ? dlopen(p, RTLD_NOW | RTLD_LOCAL) : NULL;
		if (!h) { fprintf(stderr, "recorder: set BSLV_ENGINE_SO\n"); abort(); }
	}
	return h;
}
#define NEXT(name, type) static type fn; if (!fn) fn = (type)dlsym(engine(), name)

void poly__initialise(poly_args *a)
{
	NEXT("poly__initialise", void_fn);
	fn(a);
	fprintf(out(), "{\"ev\":\"init\",\"id\":\"%p\",\"dim\":%zu}\n", (void *)a, a->dim);
}
int poly__add_vrtx(poly_args *a)
{
	NEXT("poly__add_vrtx", int_fn);
	double hp[64];
	((void (*)(double *, int, double *))a->dualV2primalH)(a->val, (int)a->ideal, hp);
	FILE *f = out();
	fprintf(f, "{\"ev\":\"add\",\"id\":\"%p\",\"ideal\":%d,", (void *)a, (int)a->ideal);
	hexvec(f, "val", a->val, a->dim);
	fprintf(f, ",");
	hexvec(f, "hp", hp, a->dim + 1);
	int rc = fn(a);
	fprintf(f, ",\"rc\":%d}\n", rc);
	return rc;
}
int poly__intl_apprx(poly_args *a)
{
	NEXT("poly__intl_apprx", int_fn);
	/* dual slot 0 may have been patched by the caller (cone_vertenum, bslv_algs.c:338-339) */
	FILE *f = out();
	fprintf(f, "{\"ev\":\"apprx\",\"id\":\"%p\",\"slot0_ideal\":%d,", (void *)a, (int)IS_ELEM(a->dual.ideal, 0));
	hexvec(f, "slot0", a->dual.data, a->dim);
	int rc = fn(a);
	fprintf(f, ",\"rc\":%d}\n", rc);
	return rc;
}
int poly__get_vrtx(poly_args *a)
{
	NEXT("poly__get_vrtx", int_fn);
	int rc = fn(a);
	fprintf(out(), "{\"ev\":\"get\",\"id\":\"%p\",\"rc\":%d,\"idx\":%zu}\n", (void *)a, rc, rc ? (size_t)0 : a->idx);
	return rc;
}
void poly__kill(poly_args *a)
{
	NEXT("poly__kill", void_fn);
	size_t pts = 0, dirs = 0;
	for (size_t s = 0; s < a->primal.cnt; s++)
		if (IS_ELEM(a->primal.used, s)) { if (IS_ELEM(a->primal.ideal, s)) dirs++; else pts++; }
	fprintf(out(), "{\"ev\":\"kill\",\"id\":\"%p\",\"points\":%zu,\"dirs\":%zu,\"slots\":%zu,\"dual_slots\":%zu}\n", (void *)a, pts, dirs, a->primal.cnt, a->dual.cnt);
	fflush(out());
	fn(a);
}
