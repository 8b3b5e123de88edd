// CutEngine: device memory management and the stream-ordered kernel pipeline of one cut.
//
// Built two ways:
//   * nvcc, sm_100a  -> the product (libbslv_poly_b200.so).  No CPU path: without a CUDA device
//                       every entry point fails loudly.
//   * g++ -DB200_EMULATE -> tests/_emul/libbslv_poly_emul.so, a host-side test double that runs
//                       the same per-item stage bodies serially so that the data-layout logic
//                       can be checked on a machine without a GPU.  Never shipped, never loaded
//                       by the product.
#include "cut_engine.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <stdexcept>

#ifndef B200_EMULATE
#include "cut_kernels.cuh"
#include "wave_kernels.cuh"
#else
#include "cut_bodies.h"
#include "wave_bodies.h"
#endif

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;
void b200_set_error(const std::string &msg) { g_last_error = msg; }
const char *b200_get_error() { return g_last_error.c_str(); }

static unsigned env_u32_early(const char *name, unsigned dflt)
{
	const char *e = getenv(name);
	return e ? (unsigned)strtoul(e, nullptr, 10) : dflt;
}
static double g_t_malloc = 0, g_t_memset = 0, g_t_free = 0;   // B200_PHASES report (us, process-wide)
static inline double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

[[noreturn]] static void fail(const std::string &msg)
{
	b200_set_error(msg);
	throw std::runtime_error(msg);
}

// ------------------------------------------------------------------ memory layer
#ifndef B200_EMULATE
#define CK(call)                                                                                      \
	do {                                                                                              \
		cudaError_t e__ = (call);                                                                     \
		if (e__ != cudaSuccess)                                                                       \
			fail(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" +    \
			     std::to_string(__LINE__) + " in " #call);                                            \
	} while (0)

static int g_device = -1;
static int g_pdl = 1;            // programmatic dependent launch between the kernels of one cut (env B200_PDL=0 disables)
static int g_tail_ctas = 16;     // cluster size of the tail kernels (env B200_TAIL_CTAS = 4, 8 or 16; 16 falls back to 8 if the device cannot host it)
static int g_k1_grid = 0;        // K1 grid policy (env B200_K1_GRID): 0 balanced rounds, 1 one block per group, 2 capped at residency
static int g_k1_it = 1;     // tile iterations whose loads a K1 thread keeps in flight (env B200_K1_IT = 1, 2 or 4)
int b200_num_devices()
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}
int b200_select_device(int dev)
{
	if (dev < 0 || dev >= b200_num_devices()) { b200_set_error("b200_set_device: no such CUDA device"); return 1; }
	g_device = dev;
	return cudaSetDevice(dev) == cudaSuccess ? 0 : 1;
}
static void bind_device()
{
	if (b200_num_devices() <= 0)
		fail("bensolve_b200: no CUDA device visible -- the cut step has no CPU fallback");
	if (g_device < 0) {
		const char *lr = getenv("LOCAL_RANK");
		int dev = lr ? atoi(lr) % b200_num_devices() : 0;
		const char *ov = getenv("B200_DEVICE");
		if (ov) dev = atoi(ov);
		g_device = dev;
	}
	CK(cudaSetDevice(g_device));
}
#define STREAM ((cudaStream_t)stream_)
// (cudaMallocAsync with the release threshold lifted was tried for these: growing arrays through the stream-ordered
// pool made the compaction's shadow allocation 5-10x slower -- the pool remaps physical memory to build large
// blocks -- so storage is plain cudaMalloc, and growth is kept rare instead: see grow_to().)
static void *dalloc(size_t bytes)
{
	void *p = nullptr;
	const double t0 = now_us();
	CK(cudaMalloc(&p, bytes ? bytes : 16));
	const double t1 = now_us();
	CK(cudaMemset(p, 0, bytes ? bytes : 16));
	g_t_malloc += t1 - t0;
	g_t_memset += now_us() - t1;
	return p;
}
static void dfree(void *p)
{
	if (!p) return;
	const double t0 = now_us();
	cudaFree(p);
	g_t_free += now_us() - t0;
}
static void d2d(void *dst, const void *src, size_t n) { if (n) CK(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice)); }
static void h2d(void *dst, const void *src, size_t n) { if (n) CK(cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice)); }
static void d2h(void *dst, const void *src, size_t n) { if (n) CK(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost)); }
#else
int b200_num_devices() { return 0; }
int b200_select_device(int) { return 0; }
static void bind_device() {}
static void *dalloc(size_t bytes) { return calloc(1, bytes ? bytes : 16); }
static void dfree(void *p) { free(p); }
static void d2d(void *dst, const void *src, size_t n) { if (n) memcpy(dst, src, n); }
static void h2d(void *dst, const void *src, size_t n) { if (n) memcpy(dst, src, n); }
static void d2h(void *dst, const void *src, size_t n) { if (n) memcpy(dst, src, n); }
#endif


// ------------------------------------------------------------------ multi-GPU communicator (one per process)
// NCCL is resolved at run time from the library the host process already loaded (torch bundles
// libnccl.so.2), so the cut engine has no link-time dependency on it.
struct Comm {
	int rank = 0, nranks = 1;
	void *nccl_comm = nullptr;
	b200_allgather_fn callback = nullptr;     // host all-gather (test double only)
};
static Comm g_comm;
int b200_comm_rank() { return g_comm.rank; }
int b200_comm_size() { return g_comm.nranks; }

#ifndef B200_EMULATE
#include <dlfcn.h>
typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*nccl_get_unique_id_t)(nccl_unique_id *);
typedef int (*nccl_comm_init_rank_t)(void **, int, nccl_unique_id, int);
typedef int (*nccl_all_gather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void *);
typedef const char *(*nccl_get_error_string_t)(int);
static struct {
	void *handle = nullptr;
	nccl_get_unique_id_t get_unique_id = nullptr;
	nccl_comm_init_rank_t comm_init_rank = nullptr;
	nccl_all_gather_t all_gather = nullptr;
	nccl_comm_destroy_t comm_destroy = nullptr;
	nccl_get_error_string_t get_error_string = nullptr;
} g_nccl;
static void load_nccl()
{
	if (g_nccl.handle) return;
	const char *names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char *n : names) {
		if (!n) continue;
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) break;
	}
	if (!g_nccl.handle) fail("bensolve_b200: cannot load libnccl.so.2 (import torch first, or set B200_NCCL_LIB)");
	g_nccl.get_unique_id = (nccl_get_unique_id_t)dlsym(g_nccl.handle, "ncclGetUniqueId");
	g_nccl.comm_init_rank = (nccl_comm_init_rank_t)dlsym(g_nccl.handle, "ncclCommInitRank");
	g_nccl.all_gather = (nccl_all_gather_t)dlsym(g_nccl.handle, "ncclAllGather");
	g_nccl.comm_destroy = (nccl_comm_destroy_t)dlsym(g_nccl.handle, "ncclCommDestroy");
	g_nccl.get_error_string = (nccl_get_error_string_t)dlsym(g_nccl.handle, "ncclGetErrorString");
	if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_gather || !g_nccl.comm_destroy) fail("bensolve_b200: libnccl lacks a required symbol");
}
static void nccl_check(int rc, const char *what)
{
	if (rc != 0) fail(std::string("NCCL error in ") + what + ": " + (g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?"));
}
int b200_comm_make_id(char out[128])
{
	try {
		load_nccl();
		nccl_unique_id id;
		nccl_check(g_nccl.get_unique_id(&id), "ncclGetUniqueId");
		memcpy(out, id.internal, 128);
		return 0;
	} catch (const std::exception &) { return 1; }
}
int b200_comm_start(int rank, int nranks, const char id_bytes[128])
{
	try {
		if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) fail("b200_comm_init: bad rank / size");
		if (nranks > 1) {
			load_nccl();
			bind_device();
			nccl_unique_id id;
			memcpy(id.internal, id_bytes, 128);
			nccl_check(g_nccl.comm_init_rank(&g_comm.nccl_comm, nranks, id, rank), "ncclCommInitRank");
		}
		g_comm.rank = rank;
		g_comm.nranks = nranks;
		return 0;
	} catch (const std::exception &) { return 1; }
}
void b200_comm_stop()
{
	if (g_comm.nccl_comm) g_nccl.comm_destroy(g_comm.nccl_comm);
	g_comm = Comm();
}
int b200_comm_set_callback(b200_allgather_fn) { b200_set_error("the product library exchanges over NCCL only"); return 1; }
#else
int b200_comm_make_id(char out[128]) { memset(out, 0, 128); return 0; }
int b200_comm_start(int rank, int nranks, const char *)
{
	if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return 1;
	g_comm.rank = rank;
	g_comm.nranks = nranks;
	return 0;
}
void b200_comm_stop() { g_comm = Comm(); }
int b200_comm_set_callback(b200_allgather_fn fn) { g_comm.callback = fn; return 0; }
#endif

struct GrowTimer {          // host time spent growing device storage (reported with B200_PHASES)
	double *acc, t0;
	const char *what;
	GrowTimer(double *a, const char *w) : acc(a), t0(now_us()), what(w) {}
	~GrowTimer()
	{
		const double dt = now_us() - t0;
		acc[8] += dt; acc[9] += 1;
		static const bool verbose = getenv("B200_PHASES") != nullptr;
		if (verbose && dt > 2000) fprintf(stderr, "[b200] growth step %s: %.1f ms (process totals so far: cudaMalloc %.1f, cudaMemset %.1f, cudaFree %.1f ms)\n", what, dt / 1e3, g_t_malloc / 1e3, g_t_memset / 1e3, g_t_free / 1e3);
	}
};
template <class T> static void regrow(T *&p, size_t new_n, size_t keep_n)
{
	T *q = (T *)dalloc(new_n * sizeof(T));
	if (p && keep_n) d2d(q, p, keep_n * sizeof(T));
	dfree(p);
	p = q;
}

static u32 round_up(u64 x, u32 m) { return (u32)(((x + m - 1) / m) * m); }
// New capacity of a growing array.  A growth step costs a cudaMalloc (0.3 ms whatever the size) + copy + cudaFree
// (0.3 ms, device-synchronising) PER ARRAY, so it is the number of steps that matters: arrays below `quad_below`
// entries grow 4x, larger ones 2x.  32-bit indices.
static u32 grow_to(u64 need, u64 have, u64 quad_below)
{
	const u64 f = have < quad_below ? 4 : 2;
	return (u32)std::min<u64>(0xFFFF0000ull, std::max<u64>(need, have * f));
}

// ------------------------------------------------------------------ multi-GPU look-ahead exchange area (one per process)
// Receive area + flag words of this rank, exported with cudaIpc and mapped by every peer (and theirs by us); the
// handles travel through one NCCL all-gather when the first device-resident batch of the process starts.  The pass
// counter is process-wide and monotone (engines run one after the other on the calling thread, every rank issues the
// same calls), so a flag word left by an earlier polytope never looks like a future pass.
struct XArea {
	bool tried = false, ok = false;
	unsigned long long *send = nullptr, *recv = nullptr;
	u32 *flag = nullptr;
	unsigned long long *peer_recv[B200_X_MAXRANKS] = {nullptr};
	u32 *peer_flag[B200_X_MAXRANKS] = {nullptr};
	u32 seq = 0;
	// the same set for the adjacent pairs of a sharded pair test
	unsigned long long *ksend = nullptr, *krecv = nullptr;
	u32 *kflag = nullptr;
	unsigned long long *kpeer_recv[B200_X_MAXRANKS] = {nullptr};
	u32 *kpeer_flag[B200_X_MAXRANKS] = {nullptr};
	u32 seq_k = 0;
};
static XArea g_x;
#ifndef B200_EMULATE
static void xarea_setup(cudaStream_t st)
{
	if (g_x.tried) return;
	g_x.tried = true;
	const int G = g_comm.nranks, me = g_comm.rank;
	if (G < 2 || G > B200_X_MAXRANKS || !g_comm.nccl_comm) return;
	if (const char *e = getenv("B200_WAVE_SHARD")) if (atoi(e) == 0) return;
	// every rank must take part in the all-gather below whatever happens locally: collect the local verdict first
	struct Rec { cudaIpcMemHandle_t recv, flag, krecv, kflag; int ok; int pad[15]; } mine, *all = nullptr;
	static_assert(sizeof(Rec) == 320, "exchange record");
	memset(&mine, 0, sizeof mine);
	const size_t recv_bytes = (size_t)G * 2 * B200_X_WORDS * 8, krecv_bytes = (size_t)G * 2 * B200_XK_WORDS * 8;
	bool ok = cudaMalloc((void **)&g_x.recv, recv_bytes) == cudaSuccess && cudaMalloc((void **)&g_x.flag, (size_t)G * 128) == cudaSuccess &&
	          cudaMalloc((void **)&g_x.send, (size_t)B200_X_WORDS * 8) == cudaSuccess &&
	          cudaMalloc((void **)&g_x.krecv, krecv_bytes) == cudaSuccess && cudaMalloc((void **)&g_x.kflag, (size_t)G * 128) == cudaSuccess &&
	          cudaMalloc((void **)&g_x.ksend, (size_t)B200_XK_WORDS * 8) == cudaSuccess;
	if (ok) {
		cudaMemset(g_x.recv, 0, recv_bytes);
		cudaMemset(g_x.flag, 0, (size_t)G * 128);
		cudaMemset(g_x.send, 0, (size_t)B200_X_WORDS * 8);
		cudaMemset(g_x.krecv, 0, krecv_bytes);
		cudaMemset(g_x.kflag, 0, (size_t)G * 128);
		cudaMemset(g_x.ksend, 0, (size_t)B200_XK_WORDS * 8);
		ok = cudaIpcGetMemHandle(&mine.recv, g_x.recv) == cudaSuccess && cudaIpcGetMemHandle(&mine.flag, g_x.flag) == cudaSuccess &&
		     cudaIpcGetMemHandle(&mine.krecv, g_x.krecv) == cudaSuccess && cudaIpcGetMemHandle(&mine.kflag, g_x.kflag) == cudaSuccess;
	}
	(void)cudaGetLastError();
	mine.ok = ok ? 1 : 0;
	Rec *d_send = (Rec *)dalloc(sizeof(Rec)), *d_recv = (Rec *)dalloc(sizeof(Rec) * G);
	h2d(d_send, &mine, sizeof mine);
	nccl_check(g_nccl.all_gather(d_send, d_recv, sizeof(Rec), 0 /* ncclChar */, g_comm.nccl_comm, st), "ncclAllGather (ipc handles)");
	CK(cudaStreamSynchronize(st));
	std::vector<Rec> got(G);
	d2h(got.data(), d_recv, sizeof(Rec) * G);
	dfree(d_send); dfree(d_recv);
	all = got.data();
	for (int g = 0; g < G; g++) ok = ok && all[g].ok;
	// opening may fail on one rank only (no peer access between two devices): a second all-gather settles the verdict
	if (ok) {
		for (int g = 0; g < G && ok; g++) {
			if (g == me) continue;
			void *a = nullptr, *b = nullptr, *c = nullptr, *e = nullptr;
			ok = cudaIpcOpenMemHandle(&a, all[g].recv, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
			     cudaIpcOpenMemHandle(&b, all[g].flag, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
			     cudaIpcOpenMemHandle(&c, all[g].krecv, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
			     cudaIpcOpenMemHandle(&e, all[g].kflag, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
			g_x.peer_recv[g] = (unsigned long long *)a;
			g_x.peer_flag[g] = (u32 *)b;
			g_x.kpeer_recv[g] = (unsigned long long *)c;
			g_x.kpeer_flag[g] = (u32 *)e;
		}
		(void)cudaGetLastError();
	}
	int *d_v = (int *)dalloc(64), *d_vs = (int *)dalloc((size_t)64 * G);
	int verdict[16] = {ok ? 1 : 0};
	h2d(d_v, verdict, 64);
	nccl_check(g_nccl.all_gather(d_v, d_vs, 64, 0, g_comm.nccl_comm, st), "ncclAllGather (ipc verdict)");
	CK(cudaStreamSynchronize(st));
	std::vector<int> vs((size_t)16 * G);
	d2h(vs.data(), d_vs, (size_t)64 * G);
	dfree(d_v); dfree(d_vs);
	for (int g = 0; g < G; g++) ok = ok && vs[(size_t)16 * g];
	g_x.ok = ok;
	if (!ok && me == 0) fprintf(stderr, "[b200] peer-mapped exchange area unavailable: look-ahead passes stay replicated on every rank\n");
}
#else
static void xarea_setup()
{
	if (g_x.tried) return;
	g_x.tried = true;
	const int G = g_comm.nranks;
	if (G < 2 || G > B200_X_MAXRANKS || !g_comm.callback) return;
	g_x.send = (unsigned long long *)calloc(B200_X_WORDS, 8);
	g_x.recv = (unsigned long long *)calloc((size_t)G * B200_X_WORDS, 8);
	g_x.ksend = (unsigned long long *)calloc(B200_XK_WORDS, 8);
	g_x.krecv = (unsigned long long *)calloc((size_t)G * B200_XK_WORDS, 8);
	g_x.ok = true;
}
#endif

// ------------------------------------------------------------------ construction
CutEngine::CutEngine(int dim) : d_(dim)
{
	if (dim < 1 || dim > B200_MAXD) fail("bensolve_b200: dimension " + std::to_string(dim) + " outside 1.." + std::to_string(B200_MAXD));
	bind_device();
	memset(&S_, 0, sizeof S_);
	S_.d = dim;
#ifndef B200_EMULATE
	cudaStream_t st;
	CK(cudaStreamCreate(&st));   // blocking stream: ordered against the default-stream memset/memcpy of the memory layer
	stream_ = st;
	for (int i = 0; i < 4; i++) { cudaEvent_t e; CK(cudaEventCreate(&e)); ev_[i] = e; }
	CK(cudaMallocHost((void **)&pinned_hdr_, sizeof(CutCtl)));
	if (const char *e = getenv("B200_K1_IT")) g_k1_it = atoi(e);
	if (const char *e = getenv("B200_K1_GRID")) g_k1_grid = atoi(e);
	if (const char *e = getenv("B200_TAIL_CTAS")) g_tail_ctas = atoi(e);
	if (const char *e = getenv("B200_PDL")) g_pdl = atoi(e);
	if (g_tail_ctas == 16) {
		// 16 CTAs per cluster is the non-portable maximum: opt in, and fall back to the portable 8 when no GPC of
		// this device can host 16 CTAs of 1024 threads
		bool ok = cudaFuncSetAttribute(k_tail<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
		          cudaFuncSetAttribute(k_tail2<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
		if (ok) {
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(16);
			cfg.blockDim = dim3(TAIL_THREADS);
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension;
			at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
			cfg.attrs = at; cfg.numAttrs = 1;
			int n1 = 0, n2 = 0;
			ok = cudaOccupancyMaxActiveClusters(&n1, k_tail<16>, &cfg) == cudaSuccess && n1 > 0 &&
			     cudaOccupancyMaxActiveClusters(&n2, k_tail2<16>, &cfg) == cudaSuccess && n2 > 0;
		}
		if (!ok) { (void)cudaGetLastError(); g_tail_ctas = 8; }
	}
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, g_device));
	num_sms_ = prop.multiProcessorCount;
	if (prop.major < 10) fail("bensolve_b200: built for sm_100a (B200); device reports sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
#else
	pinned_hdr_ = (CutCtl *)calloc(1, sizeof(CutCtl));
#endif
	S_.ctl = (CutCtl *)dalloc(sizeof(CutCtl));
	S_.cur = (CutParams *)dalloc(sizeof(CutParams));
	S_.dbg = (u64 *)dalloc(32 * sizeof(u64));
	gc_totals_ = (u32 *)dalloc(4 * sizeof(u32));
	nranks_ = g_comm.nranks;
	rank_ = g_comm.rank;
	S_.xchg_send = (u32 *)dalloc((size_t)B200_XCHG_WORDS * 4);
	S_.xchg_recv = (u32 *)dalloc((size_t)B200_XCHG_WORDS * 4 * nranks_);
	S_.nplist = (u32 *)dalloc((size_t)B200_VIS_MAX * sizeof(u32));
	S_.he_off = (u32 *)dalloc((B200_VIS_MAX + 1) * sizeof(u32));
	S_.he_own = (u32 *)dalloc(B200_HE_CAP * sizeof(u32));
	S_.he_inc = (u32 *)dalloc(B200_HE_CAP * sizeof(u32));
	S_.he_k = (u32 *)dalloc(B200_HE_CAP * sizeof(u32));
	S_.he_rank = (u32 *)dalloc(B200_HE_CAP * sizeof(u32));
	S_.he_incpre = (u32 *)dalloc(B200_HE_CAP * sizeof(u32));
	S_.he_flag = (u8 *)dalloc(B200_HE_CAP);
	S_.zmask = (u64 *)dalloc((size_t)B200_VIS_MAX * (B200_MAXINC / 64) * sizeof(u64));
	S_.cap_he = B200_HE_CAP;
	if (const char *e = getenv("B200_HE_CAP")) S_.cap_he = std::min<u32>(B200_HE_CAP, (u32)std::max(1, atoi(e)));   // test hook
	tiny_caps_ = getenv("B200_TINY_CAPS") != nullptr;
	if (tiny_caps_) {                        // test hook: start so small that every capacity negotiation path runs
		ensure_rows(B200_TILE);
		ensure_inc(64);
		ensure_adj(64);
		ensure_padj(16);
		ensure_pairs(16);
		ensure_bits(16);
		ensure_stage(B200_STAGE_HDR + 64);
		ensure_facets(8);
		return;
	}
	ensure_rows(1u << 17);               // (a few tens of MB in all: nothing on a 180 GB device, and ten growth steps saved)
	ensure_inc(1u << 20);
	ensure_adj(1u << 20);
	ensure_padj(1u << 14);
	ensure_pairs(1u << 16);
	ensure_bits(1u << 16);
	ensure_stage(1u << 18);
	ensure_facets(1024);
}

CutEngine::~CutEngine()
{
	if (getenv("B200_PHASES")) {
		fprintf(stderr, "[b200] tail phase ns:");
		for (int k = 0; k < 16; k++) fprintf(stderr, " p%d=%.1fus", k, stats_.phase_ns[k] / 1e3 / std::max<u64>(1, stats_.cuts));
		fprintf(stderr, " (per cut, %llu cuts)\n", (unsigned long long)stats_.cuts);
		fprintf(stderr, "[b200] sub-phases (thread 0 of CTA 0):");
		for (int k = 0; k < 14; k++) fprintf(stderr, " s%d=%.1fus", k, stats_.sub_ns[k] / 1e3 / std::max<u64>(1, stats_.cuts));
		fprintf(stderr, "\n");
		fprintf(stderr, "[b200] host us per cut: launch=%.1f wait=%.1f redo=%.1f unpack+gc=%.1f total_in_cut=%.1f redo_loops=%llu compactions=%llu (host ms in compaction %.1f, of which shadow allocation %.1f); device growth %.1f ms in %.0f events\n", stats_.host_us[0] / std::max<u64>(1, stats_.cuts), stats_.host_us[1] / std::max<u64>(1, stats_.cuts), stats_.host_us[2] / std::max<u64>(1, stats_.cuts), stats_.host_us[3] / std::max<u64>(1, stats_.cuts), stats_.host_us[4] / std::max<u64>(1, stats_.cuts), (unsigned long long)stats_.redo_loops, (unsigned long long)stats_.compactions, stats_.host_us[5] / 1e3, stats_.host_us[6] / 1e3, stats_.host_us[8] / 1e3, stats_.host_us[9]);
	}
#ifndef B200_EMULATE
	cudaSetDevice(g_device);
	if (stream_) cudaStreamSynchronize(STREAM);
#endif
	void *ptrs[] = {S_.coord, S_.row_slot, S_.root, flush_buf_, S_.live, S_.ideal, S_.inc_off, S_.inc_len, S_.adj_off, S_.adj_len,
	                S_.inc_pool, S_.adj_pool, S_.facet_cnt, S_.facet_alive, row_scratch_,
	                S_.padj, S_.pair_a, S_.pair_b, S_.surv_a, S_.surv_b, S_.facet_epoch, S_.facet_local, S_.bits,
#ifdef B200_EMULATE
	                S_.stage,
#endif
	                S_.nplist, S_.dbg, S_.xchg_send, S_.xchg_recv, S_.he_off, S_.he_own, S_.he_inc, S_.he_k, S_.he_rank, S_.he_incpre, S_.he_flag, S_.zmask, S_.zlong, S_.dead_facets, S_.ctl, S_.cur};
	for (void *p : ptrs) dfree(p);
	dfree(gc_totals_);
	drop_shadow();
	wave_free();
#ifndef B200_EMULATE
	for (int i = 0; i < 4; i++) if (ev_[i]) cudaEventDestroy((cudaEvent_t)ev_[i]);
	if (pinned_hdr_) cudaFreeHost(pinned_hdr_);
	if (pinned_stage_) cudaFreeHost(pinned_stage_);
	if (stream_) cudaStreamDestroy(STREAM);
#else
	free(pinned_hdr_);
	free(pinned_stage_);
	free(pinned_bulk_);
#endif
}

// ------------------------------------------------------------------ capacity
void CutEngine::drop_shadow()
{
	if (!shadow_valid_) return;
	for (int k = 0; k < 11; k++) { dfree(shadow_[k]); shadow_[k] = nullptr; }
	shadow_valid_ = false;
}

void CutEngine::ensure_rows(u32 need)
{
	if (need <= S_.cap_rows) return;
	GrowTimer gt(stats_.host_us, "ensure_rows");
	// the multi-GPU exchange packs (row | class << 30) into one word (k_xchg_pack / k_xchg_merge)
	if (nranks_ > 1 && need > (1u << 30)) fail("bensolve_b200: more than 2^30 rows are not supported with several ranks");
	drop_shadow();
	const u32 old = S_.cap_rows, keep = hdr_.nrows;
	const u32 cap = round_up(grow_to(need, old, 1ull << 21), B200_TILE);
#ifndef B200_EMULATE
	if (stream_) CK(cudaStreamSynchronize(STREAM));
#endif
	double *nc = (double *)dalloc((size_t)cap * d_ * sizeof(double));
	for (int j = 0; j < d_ && S_.coord; j++) d2d(nc + (size_t)j * cap, S_.coord + (size_t)j * old, (size_t)keep * sizeof(double));
	dfree(S_.coord);
	S_.coord = nc;
	regrow(S_.row_slot, cap, keep);
	regrow(S_.root, cap, keep);
	regrow(S_.live, cap / 32, old / 32);
	regrow(S_.ideal, cap / 32, old / 32);
	regrow(S_.inc_off, cap, keep);
	regrow(S_.inc_len, cap, keep);
	regrow(S_.adj_off, cap, keep);
	regrow(S_.adj_len, cap, keep);
	// the per-cut scratch arrays indexed by row hold nothing across a growth step: one block for all of them
	// (13 arrays = 13 x (cudaMalloc + cudaFree) otherwise), carved at 256-byte boundaries
	{
		S_.cap_tiles = cap / B200_TILE;
		const size_t c4 = (((size_t)cap * 4 + 255) & ~(size_t)255), c1 = (((size_t)cap + 255) & ~(size_t)255), t4 = (((size_t)S_.cap_tiles * 4 + 255) & ~(size_t)255);
		dfree(row_scratch_);
		row_scratch_ = dalloc(c1 + 14 * c4 + 2 * t4);
		char *q = (char *)row_scratch_;
		auto take = [&](size_t bytes) { char *r = q; q += bytes; return r; };
		S_.cls = (u8 *)take(c1);
		S_.vis = (u32 *)take(c4);
		S_.cnt3 = (u32 *)take(3 * c4);
		S_.base3 = (u32 *)take(3 * c4);
		S_.new_padj_off = (u32 *)take(c4);
		S_.new_padj_len = (u32 *)take(c4);
		S_.new_parent = (u32 *)take(c4);
		S_.deg = (u32 *)take(c4);
		S_.adj_fill = (u32 *)take(c4);
		S_.adj_base = (u32 *)take(c4);
		S_.dead_slots = (u32 *)take(c4);
		S_.tile_cnt = (u32 *)take(t4);
		S_.tile_base = (u32 *)take(t4);
	}
	small_dirty_ = true;
	S_.cap_rows = cap;
}
void CutEngine::ensure_inc(u32 need)
{
	if (need <= S_.cap_inc) return;
	GrowTimer gt(stats_.host_us, "ensure_inc");
	drop_shadow();
	u32 cap = grow_to(need, S_.cap_inc, 1ull << 24);
	regrow(S_.inc_pool, cap, hdr_.inc_used);
	S_.cap_inc = cap;
}
void CutEngine::ensure_adj(u32 need)
{
	if (need <= S_.cap_adj) return;
	GrowTimer gt(stats_.host_us, "ensure_adj");
	drop_shadow();
	u32 cap = grow_to(need, S_.cap_adj, 1ull << 24);
	regrow(S_.adj_pool, cap, hdr_.adj_used);
	S_.cap_adj = cap;
}
void CutEngine::ensure_padj(u32 need)
{
	if (need <= S_.cap_padj) return;
	GrowTimer gt(stats_.host_us, "ensure_padj");
	u32 cap = (u32)std::max<u64>(need, (u64)S_.cap_padj * 2);
	regrow(S_.padj, cap, 0);
	S_.cap_padj = cap;
}
void CutEngine::ensure_pairs(u32 need)
{
	if (need <= S_.cap_pairs) return;
	GrowTimer gt(stats_.host_us, "ensure_pairs");
	u32 cap = (u32)std::max<u64>(need, (u64)S_.cap_pairs * 2);
	regrow(S_.pair_a, cap, 0);
	regrow(S_.pair_b, cap, 0);
	regrow(S_.surv_a, cap, 0);
	regrow(S_.surv_b, cap, 0);
	S_.cap_pairs = cap;
}
void CutEngine::ensure_bits(u64 need)
{
	if (need <= S_.cap_bits) return;
	GrowTimer gt(stats_.host_us, "ensure_bits");
	u64 cap = std::max<u64>(need, S_.cap_bits * 2);
	regrow(S_.bits, cap, 0);
	S_.cap_bits = cap;
}
void CutEngine::ensure_stage(u64 need)
{
	if (need <= S_.cap_stage) return;
	GrowTimer gt(stats_.host_us, "ensure_stage");
	const u64 cap = std::max<u64>(need, S_.cap_stage * 2);
#ifndef B200_EMULATE
	// mapped pinned host memory: the kernels write the delta record straight into it (no D2H copy)
	if (stream_) CK(cudaStreamSynchronize(STREAM));
	if (pinned_stage_) CK(cudaFreeHost(pinned_stage_));
	CK(cudaHostAlloc((void **)&pinned_stage_, cap, cudaHostAllocMapped));
	memset(pinned_stage_, 0, cap);
	void *dp = nullptr;
	CK(cudaHostGetDevicePointer(&dp, pinned_stage_, 0));
	S_.stage = (unsigned char *)dp;
#else
	regrow(S_.stage, cap, 0);
	free(pinned_stage_);
	pinned_stage_ = (unsigned char *)calloc(1, cap);
#endif
	S_.cap_stage = cap;
}
void CutEngine::ensure_facets(u32 need)
{
	if (need <= S_.cap_facets) return;
	GrowTimer gt(stats_.host_us, "ensure_facets");
	u32 cap = (u32)std::max<u64>(need, (u64)S_.cap_facets * 2);
	regrow(S_.facet_cnt, cap, S_.cap_facets);
	regrow(S_.facet_alive, cap, S_.cap_facets);
	regrow(S_.facet_epoch, cap, S_.cap_facets);
	regrow(S_.facet_local, cap, S_.cap_facets);
	regrow(S_.dead_facets, cap, 0);
	S_.cap_facets = cap;
	// facet bitmaps for on-plane vertices on more than B200_MAXINC facets (only polytopes with that many facets can
	// have them): one row per visited entry, as long as that stays within 256 MB -- beyond, the direct test is used
	const u32 words = (cap + 63) / 64;
	if (cap > B200_MAXINC && (u64)words * B200_VIS_MAX * 8 <= (256ull << 20)) {
		regrow(S_.zlong, (size_t)words * B200_VIS_MAX, 0);
		S_.zlong_words = words;
	} else {
		dfree(S_.zlong);
		S_.zlong = nullptr;
		S_.zlong_words = 0;
	}
}

// ------------------------------------------------------------------ initial state
void CutEngine::upload_initial(u32 n, const double *coords_aos, const u8 *ideal,
                               const std::vector<std::vector<u32>> &inc, const std::vector<std::vector<u32>> &adj,
                               u32 n_facets, const std::vector<u32> &facet_counts)
{
	ensure_rows(n + B200_TILE);
	ensure_facets(n_facets + 64);
	std::vector<double> soa((size_t)n);
	for (int j = 0; j < d_; j++) {
		for (u32 r = 0; r < n; r++) soa[r] = coords_aos[(size_t)r * d_ + j];
		h2d(S_.coord + (size_t)j * S_.cap_rows, soa.data(), n * sizeof(double));
	}
	std::vector<u32> slot(n), live((n + 31) / 32, 0), idl((n + 31) / 32, 0), ioff(n), ilen(n), aoff(n), alen(n), ipool, apool;
	for (u32 r = 0; r < n; r++) {
		slot[r] = r;
		live[r >> 5] |= 1u << (r & 31);
		if (ideal[r]) idl[r >> 5] |= 1u << (r & 31);
		std::vector<u32> l = inc[r];
		std::sort(l.begin(), l.end());
		ioff[r] = (u32)ipool.size();
		ilen[r] = (u32)l.size();
		ipool.insert(ipool.end(), l.begin(), l.end());
		aoff[r] = (u32)apool.size();
		alen[r] = (u32)adj[r].size();
		apool.insert(apool.end(), adj[r].begin(), adj[r].end());
	}
	ensure_inc((u32)ipool.size() + (1u << 16));
	ensure_adj((u32)apool.size() + (1u << 16));
	h2d(S_.row_slot, slot.data(), n * 4);
	h2d(S_.live, live.data(), live.size() * 4);
	h2d(S_.ideal, idl.data(), idl.size() * 4);
	h2d(S_.inc_off, ioff.data(), n * 4);
	h2d(S_.inc_len, ilen.data(), n * 4);
	h2d(S_.adj_off, aoff.data(), n * 4);
	h2d(S_.adj_len, alen.data(), n * 4);
	h2d(S_.inc_pool, ipool.data(), ipool.size() * 4);
	h2d(S_.adj_pool, apool.data(), apool.size() * 4);
	std::vector<u32> fc(facet_counts), fa(n_facets, 0);
	fc.resize(n_facets, 0);
	for (u32 f = 0; f < n_facets; f++) fa[f] = fc[f] > 0;
	h2d(S_.facet_cnt, fc.data(), n_facets * 4);
	h2d(S_.facet_alive, fa.data(), n_facets * 4);
	memset(&hdr_, 0, sizeof hdr_);
	hdr_.nrows = hdr_.slot_cnt = hdr_.n_live = n;
	hdr_.inc_used = (u32)ipool.size();
	hdr_.adj_used = (u32)apool.size();
	hdr_.min_strict_row = B200_NONE;
	h2d(S_.ctl, &hdr_, sizeof hdr_);
	small_dirty_ = true;
}

// ------------------------------------------------------------------ the pipeline
#ifndef B200_EMULATE
template <int D, int IT> static void launch_classify_lists_it(const DevState &S, const CutParams &P, const double *dv, const unsigned char *di, u64 vi, u32 nrows, u32 tlo, u32 thi, int grid, cudaStream_t st)
{
	const u32 nr = nrows;
	if (dv) k_classify_lists<D, true, IT><<<grid, K_THREADS, 0, st>>>(S, P, dv, di, vi, nr, tlo, thi);
	else k_classify_lists<D, false, IT><<<grid, K_THREADS, 0, st>>>(S, P, nullptr, nullptr, 0, nr, tlo, thi);
}
// grid: one block per group of IT*512 rows, capped at the number of co-resident blocks (persistent
// grid-stride loop beyond that), so the last wave is never a partial one
template <int D> static void launch_classify_lists(const DevState &S, const CutParams &P, const double *dv, const unsigned char *di, u64 vi, u32 nrows, u32 tlo, u32 thi, int num_sms, cudaStream_t st)
{
	const bool fixed = D >= 2 && D <= 8;
	const int it = fixed ? g_k1_it : 1;
	const u32 groups = (thi - tlo) * (B200_TILE / (2 * K_THREADS) / it);
	const int per_sm = it >= 4 ? 2 : it == 2 ? 4 : 8;
	// every block gets the same number of groups (+-1): no partial last wave
	const u32 resident = (u32)(num_sms * per_sm), rounds = std::max<u32>(1, (groups + resident - 1) / resident);
	int grid = (int)std::max<u32>(1, (groups + rounds - 1) / rounds);
	if (g_k1_grid == 1) grid = (int)std::max<u32>(1, groups);                 // one block per group: hardware scheduler staggers them
	if (g_k1_grid == 2) grid = (int)std::max<u32>(1, std::min<u32>(groups, resident));
	if (it >= 4) launch_classify_lists_it<D, 4>(S, P, dv, di, vi, nrows, tlo, thi, grid, st);
	else if (it == 2) launch_classify_lists_it<D, 2>(S, P, dv, di, vi, nrows, tlo, thi, grid, st);
	else launch_classify_lists_it<D, 1>(S, P, dv, di, vi, nrows, tlo, thi, grid, st);
}
template <int D> static void launch_classify(const DevState &S, int grid, cudaStream_t st) { k_classify<D><<<grid, K_THREADS, 0, st>>>(S); }

void CutEngine::launch_classify_dim(int gcls)
{
	switch (d_) {
	case 2: launch_classify<2>(S_, gcls, STREAM); break;
	case 3: launch_classify<3>(S_, gcls, STREAM); break;
	case 4: launch_classify<4>(S_, gcls, STREAM); break;
	case 5: launch_classify<5>(S_, gcls, STREAM); break;
	case 6: launch_classify<6>(S_, gcls, STREAM); break;
	case 7: launch_classify<7>(S_, gcls, STREAM); break;
	case 8: launch_classify<8>(S_, gcls, STREAM); break;
	default: launch_classify<0>(S_, gcls, STREAM); break;
	}
}

void CutEngine::launch_part_a(const CutParams &P)
{
	small_dirty_ = true;
	const u32 ntiles = (hdr_.nrows + B200_TILE - 1) / B200_TILE;
	const int gmap = num_sms_ * 4;
	const int gcls = (int)std::max<u32>(1, std::min<u32>(ntiles, (u32)num_sms_ * 8));
	if (dev_vals_)
		k_begin_dev<<<1, 32, 0, STREAM>>>(S_, dev_vals_, dev_ideal_, dev_index_, P.facet, P.batch_first, P.seq, P.zp_done);
	else
		k_begin<<<1, 32, 0, STREAM>>>(S_, P);
	if (flags_ & 1) CK(cudaEventRecord((cudaEvent_t)ev_[0], STREAM));
	launch_classify_dim(gcls);
	if (flags_ & 1) CK(cudaEventRecord((cudaEvent_t)ev_[1], STREAM));
	k_scan_tiles<<<1, SCAN_THREADS, 0, STREAM>>>(S_);
	k_scatter<<<gcls, K_THREADS, 0, STREAM>>>(S_);
	k_zp_closure<<<1, SCAN_THREADS, 0, STREAM>>>(S_);
	k_count<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k_scan3_plan<<<1, SCAN_THREADS, 0, STREAM>>>(S_);
	k_emit<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k_dead_facets<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	stats_.kernel_launches += 9;
}

void CutEngine::launch_part_b(bool rerun)
{
	const int gmap = num_sms_ * 4;
	if (rerun) { k_pairs_reset<<<gmap, K_THREADS, 0, STREAM>>>(S_); stats_.kernel_launches++; }
	k4_assign<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k4_plan_kernel<<<1, 32, 0, STREAM>>>(S_);
	k4_zero<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k4_build<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k4_filter<<<num_sms_ * 4, K_THREADS, 0, STREAM>>>(S_, k4_threshold(S_, true), 0);
	k4_contain<<<num_sms_ * 4, K_THREADS, 0, STREAM>>>(S_);
	k_adj_scan<<<1, SCAN_THREADS, 0, STREAM>>>(S_);
	k_adj_place<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k_adj_pair_fill<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k_adj_sort<<<gmap, K_THREADS, 0, STREAM>>>(S_);
	k_finish<<<1, 32, 0, STREAM>>>(S_);
	stats_.kernel_launches += 11;
	CK(cudaGetLastError());
}

void CutEngine::launch_k1_lists(const CutParams &P, const double *dv, const unsigned char *di, u64 vi, bool sharded)
{
	u32 lo, hi;
	tile_range(sharded, lo, hi);
	{   // launched even for an empty share: block 0 publishes the halfspace on this rank's device
		switch (d_) {
		case 2: launch_classify_lists<2>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 3: launch_classify_lists<3>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 4: launch_classify_lists<4>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 5: launch_classify_lists<5>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 6: launch_classify_lists<6>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 7: launch_classify_lists<7>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		case 8: launch_classify_lists<8>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		default: launch_classify_lists<0>(S_, P, dv, di, vi, hdr_.nrows, lo, hi, num_sms_, STREAM); break;
		}
	}
	if (sharded && nranks_ > 1) {        // exchange: pack -> all-gather over NVLink -> merge
		stats_.sharded_cuts++;
		k_xchg_pack<<<1, TAIL_THREADS, 0, STREAM>>>(S_);
		nccl_check(g_nccl.all_gather(S_.xchg_send, S_.xchg_recv, (size_t)B200_XCHG_WORDS * 4, 0 /* ncclChar */, g_comm.nccl_comm, STREAM), "ncclAllGather");
		k_xchg_merge<<<1, TAIL_THREADS, 0, STREAM>>>(S_, (u32)nranks_);
		stats_.kernel_launches += 3;
	}
}

template <class K, class... A> static void launch_cluster(K kernel, int ctas, cudaStream_t st, A... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(ctas);
	cfg.blockDim = dim3(TAIL_THREADS);
	cfg.stream = st;
	cudaLaunchAttribute at[2];
	int na = 0;
	if (ctas > 1) {
		at[na].id = cudaLaunchAttributeClusterDimension;
		at[na].val.clusterDim.x = ctas; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
		na++;
	}
	if (g_pdl) {     // let this grid's launch overlap the tail of the previous one; the kernel waits in cudaGridDependencySynchronize()
		at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[na].val.programmaticStreamSerializationAllowed = 1;
		na++;
	}
	cfg.attrs = at; cfg.numAttrs = na;
	CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}
template <class K, class... A> static void launch_dependent(K kernel, int grid, int block, cudaStream_t st, A... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(block);
	cfg.stream = st;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
	CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}

// small-cut path: streaming K1 + single-CTA tail (+ multi-block K4 pair test for medium cuts)
void CutEngine::launch_small(const CutParams &P, int mode, bool header_only)
{
	if (small_dirty_) {
		k_reset_small<<<1, 32, 0, STREAM>>>(S_);
		small_dirty_ = false;
		stats_.kernel_launches++;
	}
	// tiny cuts run the tail in one CTA (block barriers); larger ones in an 8-CTA cluster
	static const u32 one_cta_max = getenv("B200_TAIL1_MAX_VIS") ? (u32)atoi(getenv("B200_TAIL1_MAX_VIS")) : 96u;
	static const u32 fuse_max_rows = getenv("B200_FUSE_MAX_ROWS") ? (u32)atoi(getenv("B200_FUSE_MAX_ROWS")) : 16384u;
	const bool force_wide = (flags_ & 16) != 0;     // test hook: every cut through the widest cluster + the grid-wide pair test
	const bool tiny = !force_wide && expect_vis_ <= one_cta_max && (expect_m_ <= B200_K4_SMALL / 2 || mode == 1);
	// a polytope of a few thousand rows is classified inside the single-CTA tail: one launch per cut
	const bool fused = tiny && hdr_.nrows <= fuse_max_rows && !dev_vals_ && !shard_now();
	if (flags_ & 1) CK(cudaEventRecord((cudaEvent_t)ev_[0], STREAM));
	if (!fused) launch_k1_lists(P, dev_vals_, dev_ideal_, dev_index_, shard_now());
	if (flags_ & 1) CK(cudaEventRecord((cudaEvent_t)ev_[1], STREAM));
	const int mode_bits = mode | (fused ? TAIL_MODE_FUSED_K1 : 0);
	const int ho = header_only ? 1 : 0;
	if (tiny) {
		launch_cluster(k_tail<1>, 1, STREAM, S_, mode_bits, ho, P, hdr_.nrows);
	} else if (g_tail_ctas == 4) {
		launch_cluster(k_tail<4>, 4, STREAM, S_, mode_bits, ho, P, hdr_.nrows);
	} else if (g_tail_ctas == 16 && wide_cluster()) {
		launch_cluster(k_tail<16>, 16, STREAM, S_, mode_bits, ho, P, hdr_.nrows);
	} else {
		launch_cluster(k_tail<TAIL_CTAS>, TAIL_CTAS, STREAM, S_, mode_bits, ho, P, hdr_.nrows);
	}
	stats_.kernel_launches += fused ? 1 : 2;
	if (mode == 1) launch_k4_and_tail2(header_only);
	CK(cudaGetLastError());
}

// The 16-CTA cluster pays (more warps, fewer lanes each) once a cut has a few thousand half-edges or several hundred
// new rows; below that the portable 8-CTA cluster has the cheaper barriers.
bool CutEngine::wide_cluster() const { return (flags_ & 16) || expect_vis_ > 256 || expect_m_ > 512; }

void CutEngine::launch_k4_and_tail2(bool header_only)
{
	// per-call path: one extra block publishes the early header (the payload k_tail wrote is complete), so that the
	// host applies the record while the pair test and the adjacency build run
	const int early = (!header_only && on_early_ && !early_done_) ? 1 : 0;
	launch_dependent(k4_filter, num_sms_ * 4 + early, K_THREADS, STREAM, S_, k4_threshold(S_, true), early);
	launch_dependent(k4_contain, num_sms_ * 4, K_THREADS, STREAM, S_);
	const int ho = header_only ? 1 : 0;
	if (g_tail_ctas == 4) launch_cluster(k_tail2<4>, 4, STREAM, S_, ho);
	else if (g_tail_ctas == 16 && wide_cluster()) launch_cluster(k_tail2<16>, 16, STREAM, S_, ho);
	else launch_cluster(k_tail2<TAIL_CTAS>, TAIL_CTAS, STREAM, S_, ho);
	stats_.kernel_launches += 3;
}

void CutEngine::launch_part_c(bool header_only)
{
	k_pack_delta<<<header_only ? 1 : num_sms_ * 2, K_THREADS, 0, STREAM>>>(S_, header_only ? 1 : 0);
	k_publish<<<1, 32, 0, STREAM>>>(S_, seq_);
	stats_.kernel_launches++;
	stats_.kernel_launches++;
}

// One D2H copy brings the header and (almost always) the whole delta; a second one only when the
// record is longer than the speculative first chunk.
void CutEngine::fetch_delta()
{
	volatile u32 *seqp = (volatile u32 *)(pinned_stage_ + B200_STAGE_SEQ);
	volatile u32 *earlyp = (volatile u32 *)(pinned_stage_ + B200_STAGE_EARLY + B200_STAGE_SEQ);
	for (u64 spins = 1; *seqp != seq_; spins++) {
		if (on_early_ && !early_done_ && !header_only_ && *earlyp == seq_) {
			// the payload is complete and final (new rows, parents, retired slots, dead facets): hand it to the
			// caller now; what is still running on the device only builds adjacency
			__sync_synchronize();
			CutCtl eh;
			memcpy(&eh, pinned_stage_ + B200_STAGE_EARLY, sizeof(CutCtl));
			const StageLayout L = stage_layout(eh, d_);
			if (!(eh.status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE | ST_NEED_BIG)) && L.total <= S_.cap_stage) {
				CutDelta e;
				e.trigger_slot = eh.min_strict_slot;
				e.n_new = eh.n_new;
				e.first_new_slot = eh.slot_cnt;              // not committed yet on the device
				e.coords = (const double *)(pinned_stage_ + L.coords);
				e.parent_slot = (const u32 *)(pinned_stage_ + L.parent);
				e.ideal = (const u8 *)(pinned_stage_ + L.ideal);
				e.dead_slots = (const u32 *)(pinned_stage_ + L.dead_slots);
				e.n_dead_entries = eh.n_vis;
				e.dead_facets = (const u32 *)(pinned_stage_ + L.dead_facets);
				e.n_dead_facets = eh.n_dead_facets;
				early_done_ = true;
				early_n_new_ = eh.n_new;
				(*on_early_)(e);
			}
		}
		if ((spins & 0x3fff) == 0) {          // every ~16k polls make sure the stream is still healthy
			cudaError_t e = cudaStreamQuery(STREAM);
			if (e != cudaSuccess && e != cudaErrorNotReady)
				fail(std::string("CUDA error while waiting for the cut: ") + cudaGetErrorString(e));
			if (e == cudaSuccess && *seqp != seq_) {
				CK(cudaStreamSynchronize(STREAM));
				if (*seqp != seq_) fail("bensolve_b200: the device finished a cut without publishing its record");
			}
		}
	}
	__sync_synchronize();
	memcpy(&hdr_, pinned_stage_, sizeof(CutCtl));
}
#else // ---- host-side test double: same stage bodies, run serially
static CutParams emu_params(const DevState &S, const CutParams &Pin, const double *dv, const unsigned char *di, u64 vi)
{
	CutParams P = Pin;
	if (dv) {
		double hh = 0;
		for (int j = 0; j < B200_MAXD; j++) { const double v = j < S.d ? dv[vi * S.d + j] : 0.0; P.h[j] = v; hh += v * v; }
		P.alpha = (di && di[vi]) ? 0.0 : -1.0;
		for (int id = 0; id < 2; id++) { const double thr = id ? 0.0 : P.alpha; P.hi[id] = thr + 1e-9; P.mid[id] = thr + 1.0e-2 * 1e-9; P.lo[id] = thr - 1e-9; }
		P.hh = hh;
	}
	return P;
}
// K1 over rows [row_lo, row_hi) + ordered compaction (multi-rank: exchange of the per-rank records
// through the host all-gather callback, then merge) + decision; returns false when the cut ends
// here (redundant)
static bool emu_classify(DevState &S, const CutParams &P, u32 row_lo = 0, u32 row_hi = B200_NONE, int nranks = 1)
{
	CutCtl *c = S.ctl;
	*S.cur = P;
	c->status = 0;
	c->n_strict = 0;
	c->min_strict_row = c->min_strict_slot = B200_NONE;
	c->n_zp = c->n_zp_projected = c->n_vis = c->n_new = c->inc_new = c->padj_new = 0;
	c->n_minus = c->n_zero = c->n_pairs = c->adj_new = c->n_dead_facets = c->n_live_scanned = 0;
	c->n_local = c->wl = c->mpad = c->n_surv = 0;
	S.facet_cnt[P.facet] = 0;
	S.facet_alive[P.facet] = 1;
	row_hi = std::min<u32>(row_hi, c->nrows);
	for (u32 r = row_lo; r < row_hi; r++) {
		bool strict, zp;
		u8 cl = classify_row(S, P, r, strict, zp);
		if (cl == CLS_DEAD) continue;
		if (cl != CLS_PLUS) S.cls[r] = cl;           // PLUS rows already read PLUS (invariant between cuts)
		else if (S.cls[r] != CLS_PLUS) fail("class invariant broken: live row does not read PLUS between cuts");
		c->n_live_scanned++;
		if (strict) { c->n_strict++; if (r < c->min_strict_row) c->min_strict_row = r; }
		if (zp) c->n_zp++;
		if (cl != CLS_PLUS) S.vis[c->n_vis++] = r;
	}
	if (nranks > 1) {
		if (!g_comm.callback) fail("multi-rank test double needs b200_comm_set_callback");
		std::vector<u32> send(B200_XCHG_WORDS, 0), recv((size_t)B200_XCHG_WORDS * nranks, 0);
		send[0] = c->n_strict; send[1] = c->min_strict_row; send[2] = c->n_zp;
		send[3] = c->n_vis > B200_XCHG_CAP ? B200_NONE : c->n_vis;
		for (u32 i = 0; i < c->n_vis && i < B200_XCHG_CAP; i++) send[4 + i] = S.vis[i] | ((u32)S.cls[S.vis[i]] << 30);
		g_comm.callback(send.data(), recv.data(), (size_t)B200_XCHG_WORDS * 4);
		c->n_strict = c->n_zp = c->n_vis = 0;
		c->min_strict_row = B200_NONE;
		bool over = false;
		for (int g = 0; g < nranks; g++) {
			const u32 *r = recv.data() + (size_t)g * B200_XCHG_WORDS;
			c->n_strict += r[0];
			c->min_strict_row = std::min(c->min_strict_row, r[1]);
			c->n_zp += r[2];
			if (r[3] == B200_NONE) { over = true; continue; }
			for (u32 i = 0; i < r[3]; i++) {
				const u32 row = r[4 + i] & 0x3FFFFFFFu;
				S.vis[c->n_vis++] = row;
				S.cls[row] = (u8)(r[4 + i] >> 30);
			}
		}
		if (over) {            // as on the device: rerun unsharded through the multi-kernel path
			if (c->n_strict) { c->status |= ST_NEED_BIG; c->min_strict_slot = S.row_slot[c->min_strict_row]; return false; }
		}
	}
	if (c->n_strict == 0) {
		c->status |= ST_REDUNDANT;
		S.facet_alive[P.facet] = 0;
		for (u32 i = 0; i < c->n_vis; i++) reset_class(S, i);
		return false;
	}
	c->min_strict_slot = S.row_slot[c->min_strict_row];
	return true;
}
static void emu_zp(DevState &S, const CutParams &P)
{
	CutCtl *c = S.ctl;
	if (!c->n_zp) return;
	bool changed;
	do {
		changed = false;
		for (u32 i = 0; i < c->n_vis; i++) changed |= zp_activate(S, P, i);
	} while (changed);
}
static void emu_scan3_plan(DevState &S)
{
	CutCtl *c = S.ctl;
	u32 carry[3] = {0, 0, 0};
	for (u32 i = 0; i < c->n_vis; i++)
		for (int k = 0; k < 3; k++) { S.base3[3 * (size_t)i + k] = carry[k]; carry[k] += S.cnt3[3 * (size_t)i + k]; }
	c->n_new = carry[0]; c->inc_new = carry[1]; c->padj_new = carry[2];
	S.facet_cnt[S.cur->facet] = carry[0];
	if ((u64)c->nrows + carry[0] > S.cap_rows) c->status |= ST_OVF_ROWS;
	if ((u64)c->inc_used + carry[1] > S.cap_inc) c->status |= ST_OVF_INC;
	if (carry[2] > S.cap_padj) c->status |= ST_OVF_PADJ;
}
static void emu_k4_matrix(DevState &S)
{
	CutCtl *c = S.ctl;
	for (u32 j = 0; j < c->n_new; j++) k4_assign_columns(S, j);
	k4_plan(S);
	if (c->status & ST_OVF_BITS) return;
	for (u64 x = 0; x < (u64)c->n_local * (c->mpad / 64); x++) k4_zero_cols(S, x);
	for (u32 j = 0; j < c->n_new; j++) k4_build_row(S, j);
}
static void emu_k4_pairs(DevState &S)
{
	CutCtl *c = S.ctl;
	const u32 M = c->n_new;
	for (u32 a = 0; a < M; a++)
		for (u32 b = a + 1; b < M; b++) k4_filter_pair(S, a, b);
	if (c->n_surv > S.cap_pairs) return;
	for (u32 s = 0; s < c->n_surv; s++) k4_contain_pair(S, s);
}
static void emu_adjacency_commit(DevState &S)
{
	CutCtl *c = S.ctl;
	if (c->status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE)) return;
	if (c->n_pairs > S.cap_pairs || c->n_surv > S.cap_pairs) { c->status |= ST_OVF_PAIRS; return; }
	const u32 M = c->n_new;
	u32 carry = 0;
	for (u32 j = 0; j < M; j++) { S.adj_base[j] = carry; carry += S.new_padj_len[j] + S.deg[j]; }
	c->adj_new = carry;
	if ((u64)c->adj_used + carry > S.cap_adj) { c->status |= ST_OVF_ADJ; return; }
	for (u32 j = 0; j < M; j++) adj_place(S, j);
	for (u32 p = 0; p < c->n_pairs; p++) adj_pair_fill(S, p);
	for (u32 j = 0; j < M; j++) adj_sort(S, j);
	c->n_live = c->n_live + c->n_new - (c->n_minus + c->n_zero);
	c->nrows += c->n_new;
	c->slot_cnt += c->n_new;
	c->inc_used += c->inc_new;
	c->adj_used += c->adj_new;
}

void CutEngine::launch_part_a(const CutParams &Pin)
{
	DevState &S = S_;
	CutCtl *c = S.ctl;
	const CutParams P = emu_params(S, Pin, dev_vals_, dev_ideal_, dev_index_);
	if (!emu_classify(S, P)) return;
	emu_zp(S, P);
	for (u32 i = 0; i < c->n_vis; i++) count_outputs(S, i);
	if (c->status & ST_ERR_DEGENERATE) return;
	emu_scan3_plan(S);
	if (c->status & ST_OVF_A) return;
	for (u32 i = 0; i < c->n_vis; i++) emit_outputs(S, P, i);
	for (u32 i = 0; i < c->n_vis; i++) collect_dead_facets(S, i);
}
void CutEngine::launch_part_b(bool rerun)
{
	DevState &S = S_;
	CutCtl *c = S.ctl;
	if (c->status & (ST_REDUNDANT | ST_OVF_A | ST_ERR_DEGENERATE)) return;
	if (rerun) { c->n_pairs = c->n_surv = 0; c->status &= ~(u32)ST_OVF_B; for (u32 j = 0; j < c->n_new; j++) S.deg[j] = 0; }
	emu_k4_matrix(S);
	if (c->status & ST_OVF_BITS) return;
	emu_k4_pairs(S);
	emu_adjacency_commit(S);
}
void CutEngine::launch_part_c(bool header_only)
{
	CutCtl *c = S_.ctl;
	const StageLayout L = stage_layout(*c, S_.d);
	CutCtl h = *c;
	h.stage_bytes = (u32)L.total;
	h.status |= emu_extra_status_;
	emu_extra_status_ = 0;
	if (!header_only && L.total > S_.cap_stage) h.status |= ST_OVF_STAGE;
	memcpy(S_.stage, &h, sizeof h);
	if (header_only || (h.status & ST_OVF_STAGE) || (c->status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE))) return;
	const u64 n = (u64)c->n_new * S_.d + c->n_new + c->n_vis + c->n_dead_facets;
	for (u64 e = 0; e < n; e++) pack_delta_item(S_, L, e, c->nrows - c->n_new);
}
// the single-CTA tail of the small-cut path, phase by phase (half-edge-parallel stage bodies)
void CutEngine::launch_small(const CutParams &Pin, int mode, bool header_only)
{
	DevState &S = S_;
	CutCtl *c = S.ctl;
	const CutParams P = emu_params(S, Pin, dev_vals_, dev_ideal_, dev_index_);
	u32 tlo, thi;
	const bool shard = shard_now();
	tile_range(shard, tlo, thi);
	if (shard) stats_.sharded_cuts++;
	if (!emu_classify(S, P, tlo * B200_TILE, thi * B200_TILE, shard ? nranks_ : 1)) { launch_part_c(header_only); return; }
	if (c->n_vis > B200_VIS_MAX) { c->status |= ST_NEED_BIG; launch_part_c(header_only); return; }
	emu_zp(S, P);
	u32 H = 0;
	for (u32 i = 0; i < c->n_vis; i++) {
		const u32 r = S.vis[i];
		S.he_off[i] = H;
		H += is_visited_class(S.cls[r]) ? S.adj_len[r] : 0;
		for (u32 w = 0; w < (S.inc_len[r] + 63) / 64 && w < B200_MAXINC / 64; w++) S.zmask[(size_t)i * (B200_MAXINC / 64) + w] = 0;
	}
	S.he_off[c->n_vis] = H;
	if (H > S.cap_he) { c->status |= ST_NEED_BIG; launch_part_c(header_only); return; }
	for (u32 i = 0; i < c->n_vis; i++) he_owner_fill(S, i);
	for (u32 e = 0; e < H; e++) he_eval(S, e);
	for (u32 i = 0; i < c->n_vis; i++) {
		he_count(S, i);
		const u8 cl = S.cls[S.vis[i]];
		c->n_minus += (cl == CLS_MINUS);
		c->n_zero += (cl == CLS_ZERO);
	}
	emu_scan3_plan(S);
	if (c->status & (ST_OVF_A | ST_ERR_DEGENERATE)) { launch_part_c(header_only); return; }
	for (u32 e = H; e-- > 0;) he_emit(S, P, e);            // any order is valid: run it backwards here
	for (u32 i = 0; i < c->n_vis; i++) he_finish_vertex(S, P, i);
	for (u32 i = 0; i < c->n_vis; i++) collect_dead_facets(S, i);
	emu_k4_matrix(S);
	if (!(c->status & ST_OVF_BITS)) {
		if (mode == 1) { launch_k4_and_tail2(header_only); return; }
		if (c->n_new > B200_K4_SMALL) { emu_extra_status_ = ST_K4_PENDING; launch_part_c(true); return; }
		emu_k4_pairs(S);
	}
	emu_adjacency_commit(S);
	launch_part_c(header_only);
}
void CutEngine::launch_k4_and_tail2(bool header_only)
{
	if (!(S_.ctl->status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE))) emu_k4_pairs(S_);
	emu_adjacency_commit(S_);
	launch_part_c(header_only);
}
void CutEngine::fetch_delta()
{
	memcpy(&hdr_, S_.stage, sizeof(CutCtl));
	if (!header_only_ && !(hdr_.status & (ST_OVF_A | ST_OVF_B | ST_OVF_STAGE | ST_NEED_BIG | ST_K4_PENDING))) memcpy(pinned_stage_, S_.stage, hdr_.stage_bytes);
}
#endif

// Is K1 of the next cut split across the ranks?  The exchange (pack, NCCL all-gather, merge) costs three launches and
// ~25 us of collective latency per cut, K1 ~10 us per 10^6 rows: below a few million rows every rank classifies all
// rows itself and nothing is exchanged (the ranks stay bit-identical either way).
bool CutEngine::shard_now() const
{
	static const u32 min_rows = env_u32_early("B200_SHARD_MIN_ROWS", 4000000);
	return nranks_ > 1 && hdr_.nrows >= min_rows;
}
// this rank's share of the tiles (all of them on one GPU); ranges ascend with the rank
void CutEngine::tile_range(bool sharded, u32 &lo, u32 &hi) const
{
	const u32 ntiles = (hdr_.nrows + B200_TILE - 1) / B200_TILE;
	lo = 0;
	hi = ntiles;
	if (!sharded || nranks_ == 1) return;
	const u32 per = (ntiles + nranks_ - 1) / nranks_;
	lo = std::min<u32>(ntiles, per * rank_);
	hi = std::min<u32>(ntiles, lo + per);
}

bool CutEngine::use_small_path() const
{
#ifndef B200_EMULATE
	return true;
#else
	return (flags_ & 8) != 0;       // the test double runs the tail phases only when asked to
#endif
}

void CutEngine::bump_seq()
{
	++seq_;
#ifndef B200_EMULATE
	k_set_seq<<<1, 32, 0, STREAM>>>(S_, seq_);
#else
	S_.cur->seq = seq_;
#endif
}

void CutEngine::run_cut(const CutParams &P, bool header_only)
{
#ifndef B200_EMULATE
	CK(cudaSetDevice(g_device));
#endif
	header_only_ = header_only;
	ensure_facets(P.facet + 1);
	// head-room for the appends; exact needs are checked on the device before any mutation
	if (!tiny_caps_) ensure_rows(hdr_.nrows + std::max<u32>(4096, hdr_.n_live / 2 + 64));
#ifndef B200_EMULATE
	if (flags_ & 1) CK(cudaEventRecord((cudaEvent_t)ev_[2], STREAM));
#endif
	CutParams Pq = P;                        // P + the sequence number the device publishes when the record is staged
	const bool force_big = (flags_ & 4) != 0;
	bool small = use_small_path() && !force_big && !prefer_big_;
	auto launch_all = [&]() {
		Pq.seq = ++seq_;
		if (small) {
			launch_small(Pq, ((flags_ & 16) || expect_m_ > (B200_K4_SMALL * 7) / 8) ? 1 : 0, header_only);
		} else {
			launch_part_a(Pq);
			launch_part_b(false);
			launch_part_c(header_only);
		}
	};
	const double t_a = now_us();
	launch_all();
	const double t_b = now_us();
	fetch_delta();
	const double t_c = now_us();
	stats_.host_us[0] += t_b - t_a;
	stats_.host_us[1] += t_c - t_b;
	const u32 redo = ST_OVF_A | ST_OVF_B | ST_OVF_STAGE | ST_NEED_BIG | ST_K4_PENDING;
	for (int guard = 0; hdr_.status & redo; guard++) {
		if (guard > 16) fail("bensolve_b200: capacity negotiation did not converge");
		stats_.redo_loops++;
		if (hdr_.n_zp_projected) Pq.zp_done = 1;    // the attempt that bailed out projected ZERO+ rows in place: not twice
		if (hdr_.status & ST_NEED_BIG) {            // nothing was mutated: rerun through the multi-kernel path
			small = false;
			prefer_big_ = true;
			launch_all();
		} else if (hdr_.status & ST_K4_PENDING) {   // tail stopped before the pair test
			bump_seq();
			launch_k4_and_tail2(header_only);
		} else if (hdr_.status & ST_OVF_A) {
			if (hdr_.status & ST_OVF_ROWS) ensure_rows(hdr_.nrows + hdr_.n_new + B200_TILE);
			if (hdr_.status & ST_OVF_INC) ensure_inc(hdr_.inc_used + hdr_.inc_new);
			if (hdr_.status & ST_OVF_PADJ) ensure_padj(hdr_.padj_new);
			launch_all();
		} else if (hdr_.status & ST_OVF_B) {
			if (hdr_.status & ST_OVF_PAIRS) ensure_pairs(std::max(hdr_.n_pairs, hdr_.n_surv));
			if (hdr_.status & ST_OVF_BITS) ensure_bits(k4_words(hdr_.wl, hdr_.mpad, hdr_.n_local));
			if (hdr_.status & ST_OVF_ADJ) ensure_adj(hdr_.adj_used + hdr_.adj_new);
			bump_seq();
			launch_part_b(true);
			launch_part_c(header_only);
		} else {
			ensure_stage(hdr_.stage_bytes);
			bump_seq();
			launch_part_c(header_only);
		}
		fetch_delta();
	}
	expect_m_ = hdr_.n_new;
	expect_vis_ = hdr_.n_vis;
	stats_.host_us[2] += now_us() - t_c;
	if (prefer_big_ && hdr_.n_vis < B200_VIS_MAX / 4) prefer_big_ = false;
#ifndef B200_EMULATE
	if (flags_ & 1) {
		CK(cudaEventRecord((cudaEvent_t)ev_[3], STREAM));
		CK(cudaEventSynchronize((cudaEvent_t)ev_[3]));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev_[0], (cudaEvent_t)ev_[1]));
		stats_.classify_ms += ms;
		CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev_[2], (cudaEvent_t)ev_[3]));
		stats_.cut_ms += ms;
		{
			u64 tp[28] = {0};
			d2h(tp, S_.dbg, sizeof tp);
			if (tp[16] > tp[3] && tp[4] >= tp[16]) { stats_.sub_ns[0] += tp[16] - tp[3]; stats_.sub_ns[1] += tp[4] - tp[16]; }
			if (tp[17] >= tp[4] && tp[18] >= tp[17] && tp[5] >= tp[18]) { stats_.sub_ns[2] += tp[17] - tp[4]; stats_.sub_ns[3] += tp[18] - tp[17]; stats_.sub_ns[4] += tp[5] - tp[18]; }
			if (tp[11] >= tp[0] && tp[20] >= tp[11] && tp[21] >= tp[20] && tp[22] >= tp[21] && tp[23] >= tp[22] && tp[24] >= tp[23] && tp[25] >= tp[24] && tp[12] >= tp[25]) {
				stats_.sub_ns[7] += tp[20] - tp[11]; stats_.sub_ns[8] += tp[21] - tp[20]; stats_.sub_ns[9] += tp[22] - tp[21]; stats_.sub_ns[10] += tp[23] - tp[22];
				stats_.sub_ns[11] += tp[24] - tp[23]; stats_.sub_ns[12] += tp[25] - tp[24]; stats_.sub_ns[13] += tp[12] - tp[25];
			}
			if (tp[19] >= tp[5] && tp[6] >= tp[19]) { stats_.sub_ns[5] += tp[19] - tp[5]; stats_.sub_ns[6] += tp[6] - tp[19]; }
			for (int k = 0; k < 12; k++) if (tp[k] >= tp[0] && tp[k + 1] > tp[k] && k != 10) stats_.phase_ns[k] += tp[k + 1] - tp[k];
			if (tp[11] >= tp[0] && tp[7] >= tp[0] && tp[11] > tp[7] && tp[9] < tp[0]) {
				stats_.phase_ns[12] += tp[11] - tp[7];   // k4_filter + k4_contain + launch gaps
				if (tp[13] > tp[7] && tp[14] > tp[13] && tp[11] > tp[14]) { stats_.phase_ns[13] += tp[13] - tp[7]; stats_.phase_ns[14] += tp[14] - tp[13]; stats_.phase_ns[15] += tp[11] - tp[14]; }
			}
		}
		static const char *trace = getenv("B200_TRACE");
		if (trace && ms > atof(trace))
			fprintf(stderr, "[b200] slow cut %.3f ms: facet=%u nrows=%u live=%u n_vis=%u n_new=%u n_local=%u wl=%u n_surv=%u n_pairs=%u status=%u stage=%u\n",
			        ms, P.facet, hdr_.nrows, hdr_.n_live, hdr_.n_vis, hdr_.n_new, hdr_.n_local, hdr_.wl, hdr_.n_surv, hdr_.n_pairs, hdr_.status, hdr_.stage_bytes);
	}
#endif
	if (hdr_.status & ST_ERR_DEGENERATE)
		fail("bensolve_b200: a vertex on the cutting hyperplane lies on more than " + std::to_string(B200_MAXINC) + " facets");
}

// statistics of the cut just fetched (SURVEY 8(d) algorithmic bytes)
void CutEngine::account(const CutParams &P, u32 n_live_before, u32 nrows_before)
{
	stats_.vertex_evals += hdr_.n_live_scanned;
	stats_.rows_scanned += nrows_before;
	if (hdr_.status & ST_REDUNDANT) {
		stats_.redundant++;
		stats_.algorithmic_bytes += (u64)n_live_before * (8 * d_ + 1);
		return;
	}
	const u64 N = n_live_before, nm = hdr_.n_minus, nz = hdr_.n_zero, M = hdr_.n_new, E = M - nz;
	const u64 W = (P.facet + 64) / 64, A = hdr_.n_pairs;
	stats_.cuts++;
	stats_.minus += nm;
	stats_.zero += nz;
	stats_.zero_plus_projected += hdr_.n_zp_projected;
	stats_.edge_vertices += E;
	stats_.copies += nz;
	stats_.pair_tests += M * (M - (M ? 1 : 0)) / 2;
	stats_.new_adjacent_pairs += A;
	stats_.algorithmic_bytes += N * (8 * d_ + 1) + N + 4 * (nm + nz) + E * (24 * d_ + 24 * W) + nz * (16 * d_ + 16 * W) + 8 * M * W + 8 * A;
}

void CutEngine::cut(const CutParams &P, CutDelta &out, const std::function<void(const CutDelta &)> *on_early)
{
	const double t_in = now_us();
	out = CutDelta();
	const u32 n_live_before = hdr_.n_live, nrows_before = hdr_.nrows;
	on_early_ = (on_early && *on_early) ? on_early : nullptr;
	early_done_ = false;
	try {
		run_cut(P, false);
	} catch (...) {
		on_early_ = nullptr;
		throw;
	}
	on_early_ = nullptr;
	out.applied_early = early_done_;
	if (early_done_ && ((hdr_.status & ST_REDUNDANT) || hdr_.n_new != early_n_new_))
		fail("bensolve_b200: the early delta record disagrees with the final header");
	const double t_run = now_us();
	account(P, n_live_before, nrows_before);
	if (hdr_.status & ST_REDUNDANT) { out.redundant = 1; return; }
	// ---- the delta record is consumed in place (pinned staging buffer)
	const u32 n_new = hdr_.n_new;
	const StageLayout L = stage_layout(hdr_, d_);
	out.trigger_slot = hdr_.min_strict_slot;
	out.n_new = n_new;
	out.first_new_slot = hdr_.slot_cnt - n_new;
	out.coords = (const double *)(pinned_stage_ + L.coords);
	out.parent_slot = (const u32 *)(pinned_stage_ + L.parent);
	out.ideal = (const u8 *)(pinned_stage_ + L.ideal);
	out.dead_slots = (const u32 *)(pinned_stage_ + L.dead_slots);
	out.n_dead_entries = hdr_.n_vis;
	out.dead_facets = (const u32 *)(pinned_stage_ + L.dead_facets);
	out.n_dead_facets = hdr_.n_dead_facets;
	maybe_compact();
	stats_.host_us[3] += now_us() - t_run;
	stats_.host_us[4] += now_us() - t_in;
}

int CutEngine::cut_from_device(const double *d_vals, const unsigned char *d_ideal, u64 i, u32 facet, u32 batch_first)
{
	CutParams P;
	memset(&P, 0, sizeof P);
	P.facet = facet;
	P.batch_first = batch_first;
	dev_vals_ = d_vals;
	dev_ideal_ = d_ideal;
	dev_index_ = i;
	const u32 n_live_before = hdr_.n_live, nrows_before = hdr_.nrows;
	try {
		run_cut(P, true);
	} catch (...) {
		dev_vals_ = nullptr;
		throw;
	}
	dev_vals_ = nullptr;
	account(P, n_live_before, nrows_before);
	const int redundant = (hdr_.status & ST_REDUNDANT) ? 1 : 0;
	if (!redundant) maybe_compact();
	return redundant;
}

void CutEngine::reserve(u64 rows, u64 inc_entries, u64 adj_entries)
{
	if (rows > 0xFFFF0000ull || inc_entries > 0xFFFF0000ull || adj_entries > 0xFFFF0000ull) fail("bensolve_b200: reserve beyond 32-bit row/pool indices");
	ensure_rows((u32)rows);
	ensure_inc((u32)inc_entries);
	ensure_adj((u32)adj_entries);
	ensure_facets((u32)std::min<u64>(rows, 1u << 16));     // (a few MB: the facet-indexed arrays do not grow mid-run either)
	ensure_shadow();        // the compaction's second set of arrays belongs to "no allocation after reserve" too
}

// the shadow set of persistent arrays the compaction gathers into (allocated once per capacity); true if (re)allocated
bool CutEngine::ensure_shadow()
{
#ifndef B200_EMULATE
	const u32 cap = S_.cap_rows;
	if (shadow_valid_ && shadow_rows_ == cap && shadow_inc_ == S_.cap_inc && shadow_adj_ == S_.cap_adj) return false;
	CK(cudaSetDevice(g_device));
	drop_shadow();
	shadow_[0] = dalloc((size_t)cap * d_ * sizeof(double));
	for (int k = 1; k <= 2; k++) shadow_[k] = dalloc((size_t)cap * 4);          // row_slot, root
	for (int k = 3; k <= 4; k++) shadow_[k] = dalloc((size_t)cap / 8);           // live, ideal
	for (int k = 5; k <= 8; k++) shadow_[k] = dalloc((size_t)cap * 4);          // inc_off, inc_len, adj_off, adj_len
	shadow_[9] = dalloc((size_t)S_.cap_inc * 4);
	shadow_[10] = dalloc((size_t)S_.cap_adj * 4);
	shadow_rows_ = cap; shadow_inc_ = S_.cap_inc; shadow_adj_ = S_.cap_adj;
	shadow_valid_ = true;
	return true;
#else
	return false;
#endif
}

void CutEngine::download_mirror(MirrorDump &o, u32 n_facets)
{
	const u32 n = hdr_.nrows;
	const size_t nw = (n + 31) / 32;
	ensure_facets(n_facets);
	// one pinned buffer, one stream-ordered copy per array: pageable vectors cost more in page faults and
	// zero-fill than the transfer itself at 10^6 rows
	size_t off = 0;
	auto take = [&](size_t bytes) { const size_t o0 = off; off += (bytes + 255) & ~(size_t)255; return o0; };
	const size_t o_slot = take((size_t)n * 4), o_root = take((size_t)n * 4), o_live = take(nw * 4), o_ideal = take(nw * 4),
	             o_facet = take((size_t)n_facets * 4), o_coord = take((size_t)n * d_ * sizeof(double));
#ifndef B200_EMULATE
	// pinned staging shared by the engines of this host thread (grow-only): pinning ~100 MB per engine would cost
	// more than the faster copy saves, a pageable destination runs at ~2 GB/s once its page faults are counted
	static thread_local unsigned char *t_bulk = nullptr;
	static thread_local size_t t_bulk_cap = 0;
	CK(cudaSetDevice(g_device));
	if (off > t_bulk_cap) {
		const size_t cap = std::max(off, t_bulk_cap * 2);
		if (t_bulk) CK(cudaFreeHost(t_bulk));
		t_bulk = nullptr; t_bulk_cap = 0;
		CK(cudaMallocHost((void **)&t_bulk, cap));
		t_bulk_cap = cap;
	}
	pinned_bulk_ = t_bulk;
#else
	if (off > pinned_bulk_cap_) {
		free(pinned_bulk_);
		pinned_bulk_ = (unsigned char *)malloc(off);
		pinned_bulk_cap_ = off;
	}
#endif
	unsigned char *b = pinned_bulk_;
#ifndef B200_EMULATE
	CK(cudaSetDevice(g_device));
	auto get = [&](size_t o0, const void *src, size_t bytes) { if (bytes) CK(cudaMemcpyAsync(b + o0, src, bytes, cudaMemcpyDeviceToHost, STREAM)); };
#else
	auto get = [&](size_t o0, const void *src, size_t bytes) { if (bytes) memcpy(b + o0, src, bytes); };
#endif
	get(o_slot, S_.row_slot, (size_t)n * 4);
	get(o_root, S_.root, (size_t)n * 4);
	get(o_live, S_.live, nw * 4);
	get(o_ideal, S_.ideal, nw * 4);
	get(o_facet, S_.facet_alive, (size_t)n_facets * 4);
	for (int j = 0; j < d_; j++) get(o_coord + (size_t)j * n * sizeof(double), S_.coord + (size_t)j * S_.cap_rows, (size_t)n * sizeof(double));
#ifndef B200_EMULATE
	CK(cudaStreamSynchronize(STREAM));
#endif
	o.nrows = n;
	o.slot_cnt = hdr_.slot_cnt;
	o.row_slot = (const u32 *)(b + o_slot);
	o.root = (const u32 *)(b + o_root);
	o.live_words = (const u32 *)(b + o_live);
	o.ideal_words = (const u32 *)(b + o_ideal);
	o.facet_alive = (const u32 *)(b + o_facet);
	o.coords_soa = (const double *)(b + o_coord);
}

// K6: adjacency among the live facets (dual polytope), by the same AND+POPC filter and containment
// test as K4 on the transposed incidence: rows = live facets (ranked 0..M-1 through `facet_rank`),
// columns = live vertices (device rows after a compaction).  Returns pairs of facet RANKS.
void CutEngine::dual_adjacency(const std::vector<u32> &facet_rank, u32 M, std::vector<u32> &pair_a, std::vector<u32> &pair_b)
{
	compact();                                   // columns = dense live rows
	const u32 N = hdr_.nrows, wl = (N + 63) / 64, mpad = (M + 31) & ~31u;
	ensure_facets((u32)facet_rank.size());
	ensure_rows(std::max<u32>(hdr_.nrows, M) + 64);    // deg[] is indexed by facet rank here
	ensure_bits((u64)wl * mpad);
	h2d(S_.facet_local, facet_rank.data(), facet_rank.size() * 4);
	pair_a.clear();
	pair_b.clear();
	if (M < 2) return;
	for (int guard = 0;; guard++) {
		if (guard > 8) fail("bensolve_b200: dual adjacency buffers did not converge");
#ifndef B200_EMULATE
		CK(cudaSetDevice(g_device));
		CK(cudaMemsetAsync(S_.bits, 0, (size_t)wl * mpad * 8, STREAM));
		CK(cudaMemsetAsync(S_.deg, 0, (size_t)M * 4, STREAM));
		k6_begin<<<1, 32, 0, STREAM>>>(S_, M, wl, mpad);
		k6_build<<<num_sms_ * 8, K_THREADS, 0, STREAM>>>(S_, mpad);
		k4_filter<<<num_sms_ * 8, K_THREADS, 0, STREAM>>>(S_, k4_threshold(S_, false), 0);
		k6_contain<<<num_sms_ * 4, K_THREADS, 0, STREAM>>>(S_);
		stats_.kernel_launches += 4;
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(pinned_hdr_, S_.ctl, sizeof(CutCtl), cudaMemcpyDeviceToHost, STREAM));
		CK(cudaStreamSynchronize(STREAM));
		const CutCtl h = *pinned_hdr_;
#else
		memset(S_.bits, 0, (size_t)wl * mpad * 8);
		memset(S_.deg, 0, (size_t)M * 4);
		CutCtl *c = S_.ctl;
		c->status = 0; c->n_new = M; c->wl = wl; c->mpad = mpad; c->n_surv = c->n_pairs = 0;
		for (u32 r = 0; r < N; r++) k6_set_row_bits(S_, r, mpad);
		for (u32 a = 0; a < M; a++)
			for (u32 b = a + 1; b < M; b++) k4_filter_pair_in(S_, S_.bits, wl, mpad, a, b, k4_threshold(S_, false));
		if (c->n_surv <= S_.cap_pairs)
			for (u32 sv = 0; sv < c->n_surv; sv++) k6_contain_pair(S_, sv);
		const CutCtl h = *c;
#endif
		if (h.n_surv > S_.cap_pairs || h.n_pairs > S_.cap_pairs) { ensure_pairs(std::max(h.n_surv, h.n_pairs)); continue; }
		pair_a.resize(h.n_pairs);
		pair_b.resize(h.n_pairs);
		d2h(pair_a.data(), S_.pair_a, (size_t)h.n_pairs * 4);
		d2h(pair_b.data(), S_.pair_b, (size_t)h.n_pairs * 4);
		break;
	}
	// the cut path owns these control fields again
#ifndef B200_EMULATE
	CK(cudaMemcpyAsync(S_.ctl, &hdr_, sizeof(CutCtl), cudaMemcpyHostToDevice, STREAM));
	CK(cudaStreamSynchronize(STREAM));
#else
	*S_.ctl = hdr_;
#endif
	small_dirty_ = true;
}

void *CutEngine::device_alloc(size_t bytes) { return dalloc(bytes); }
void CutEngine::device_free(void *p) { dfree(p); }
void CutEngine::device_upload(void *dst, const void *src, size_t bytes) { h2d(dst, src, bytes); }
void CutEngine::device_download(void *dst, const void *src, size_t bytes)
{
#ifndef B200_EMULATE
	CK(cudaStreamSynchronize(STREAM));
#endif
	d2h(dst, src, bytes);
}

double CutEngine::classify_bench(const CutParams &P, int iters, int flush_l2)
{
#ifndef B200_EMULATE
	CK(cudaSetDevice(g_device));
	ensure_facets(P.facet + 1);
	const size_t flush_bytes = (size_t)256 << 20;     // > 126 MB L2
	if (flush_l2 && !flush_buf_) flush_buf_ = dalloc(flush_bytes);
	double total = 0;
	for (int it = 0; it < iters; it++) {
		k_reset_small<<<1, 32, 0, STREAM>>>(S_);
		if (flush_l2) k_flush_read<<<num_sms_ * 8, K_THREADS, 0, STREAM>>>((const uint4 *)flush_buf_, flush_bytes / 16, (unsigned *)S_.dbg + 60);
		CK(cudaEventRecord((cudaEvent_t)ev_[0], STREAM));
		launch_k1_lists(P, nullptr, nullptr, 0, false);     // the streaming K1 of the cut path, exactly as launch_small() launches it
		CK(cudaEventRecord((cudaEvent_t)ev_[1], STREAM));
		CK(cudaEventSynchronize((cudaEvent_t)ev_[1]));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, (cudaEvent_t)ev_[0], (cudaEvent_t)ev_[1]));
		total += ms;
	}
	// the probe halfspace is not a facet: undo what K1 registered, and have the next cut clear the
	// tile lists and trigger counters it left behind
	CK(cudaMemsetAsync(S_.facet_alive + P.facet, 0, 4, STREAM));
	CK(cudaStreamSynchronize(STREAM));
	small_dirty_ = true;
	return iters > 0 ? total / iters : 0.0;
#else
	(void)P; (void)iters; (void)flush_l2;
	return 0.0;
#endif
}

void CutEngine::maybe_compact()
{
	const bool eager = flags_ & 2;          // test hook: compact as soon as one row is dead
	if (hdr_.nrows == hdr_.n_live) return;
	if (!eager && (hdr_.nrows < 4 * B200_TILE || hdr_.nrows < 2 * hdr_.n_live)) return;
	compact();
}

#ifndef B200_EMULATE
void CutEngine::compact()
{
	CK(cudaSetDevice(g_device));
	const u32 nrows = hdr_.nrows, n_live = hdr_.n_live;
	if (nrows == n_live) return;
	const double t_gc0 = now_us();
	double t_alloc = 0;
	const u32 ntiles = (nrows + B200_TILE - 1) / B200_TILE, ltiles = std::max<u32>(1, (n_live + B200_TILE - 1) / B200_TILE);
	u32 *remap = S_.vis, *old_of = S_.dead_slots, *new_inc_off = S_.adj_base, *new_adj_off = S_.adj_fill;
	u32 *totals = gc_totals_;                 // persistent: an allocation per compaction is a device-wide sync (and a peer mapping under NCCL)
	// 1. remap = exclusive scan of the live bits
	LiveBitOf lb{S_.live};
	k_gscan_reduce<<<ntiles, K_THREADS, 0, STREAM>>>(lb, nrows, S_.tile_cnt);
	k_gscan_tiles<<<1, SCAN_THREADS, 0, STREAM>>>(S_.tile_cnt, ntiles, S_.tile_base, totals + 0);
	k_gscan_apply<<<ntiles, K_THREADS, 0, STREAM>>>(lb, nrows, S_.tile_base, remap);
	k_gc_invert<<<num_sms_ * 4, K_THREADS, 0, STREAM>>>(S_.live, remap, nrows, old_of);
	// 2. new pool offsets
	LenOfOld li{S_.inc_len, old_of}, la{S_.adj_len, old_of};
	k_gscan_reduce<<<ltiles, K_THREADS, 0, STREAM>>>(li, n_live, S_.tile_cnt);
	k_gscan_tiles<<<1, SCAN_THREADS, 0, STREAM>>>(S_.tile_cnt, ltiles, S_.tile_base, totals + 1);
	k_gscan_apply<<<ltiles, K_THREADS, 0, STREAM>>>(li, n_live, S_.tile_base, new_inc_off);
	k_gscan_reduce<<<ltiles, K_THREADS, 0, STREAM>>>(la, n_live, S_.tile_cnt);
	k_gscan_tiles<<<1, SCAN_THREADS, 0, STREAM>>>(S_.tile_cnt, ltiles, S_.tile_base, totals + 2);
	k_gscan_apply<<<ltiles, K_THREADS, 0, STREAM>>>(la, n_live, S_.tile_base, new_adj_off);
	// 3. gather into the shadow set of persistent arrays (allocated once per capacity), then swap
	const u32 cap = S_.cap_rows;
	{
		const double ta0 = now_us();
		if (!ensure_shadow()) {                // only the bitsets rely on zero fill beyond the live rows
			CK(cudaMemsetAsync(shadow_[3], 0, (size_t)cap / 8, STREAM));
			CK(cudaMemsetAsync(shadow_[4], 0, (size_t)cap / 8, STREAM));
		} else
			t_alloc = now_us() - ta0;
	}
	GcTarget T;
	T.coord = (double *)shadow_[0];
	T.row_slot = (u32 *)shadow_[1]; T.root = (u32 *)shadow_[2];
	T.live = (u32 *)shadow_[3]; T.ideal = (u32 *)shadow_[4];
	T.inc_off = (u32 *)shadow_[5]; T.inc_len = (u32 *)shadow_[6]; T.adj_off = (u32 *)shadow_[7]; T.adj_len = (u32 *)shadow_[8];
	T.inc_pool = (u32 *)shadow_[9]; T.adj_pool = (u32 *)shadow_[10];
	k_gc_gather<<<num_sms_ * 8, K_THREADS, 0, STREAM>>>(S_, T, n_live, remap, old_of, new_inc_off, new_adj_off);
	CK(cudaMemsetAsync(S_.cls, 0, nrows, STREAM));     // rows moved: every live row reads PLUS again
	k_gc_finish<<<1, 32, 0, STREAM>>>(S_, n_live, totals + 1, totals + 2);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(pinned_hdr_, S_.ctl, sizeof(CutCtl), cudaMemcpyDeviceToHost, STREAM));
	CK(cudaStreamSynchronize(STREAM));
	hdr_ = *pinned_hdr_;
	if (hdr_.nrows != n_live) fail("bensolve_b200: compaction lost rows");
	void *old[11] = {S_.coord, S_.row_slot, S_.root, S_.live, S_.ideal, S_.inc_off, S_.inc_len, S_.adj_off, S_.adj_len, S_.inc_pool, S_.adj_pool};
	for (int k = 0; k < 11; k++) shadow_[k] = old[k];      // the previous arrays become the next shadow set
	S_.coord = T.coord; S_.row_slot = T.row_slot; S_.root = T.root; S_.live = T.live; S_.ideal = T.ideal;
	S_.inc_off = T.inc_off; S_.inc_len = T.inc_len; S_.adj_off = T.adj_off; S_.adj_len = T.adj_len;
	S_.inc_pool = T.inc_pool; S_.adj_pool = T.adj_pool;
	stats_.compactions++;
	stats_.kernel_launches += 12;
	stats_.host_us[5] += now_us() - t_gc0;
	stats_.host_us[6] += t_alloc;
}
#else
void CutEngine::compact()
{
	DevState &S = S_;
	const u32 nrows = hdr_.nrows, n_live = hdr_.n_live, cap = S.cap_rows;
	if (nrows == n_live) return;
	std::vector<u32> remap(nrows, 0), old_of;
	for (u32 r = 0; r < nrows; r++) if (bit_test(S.live, r)) { remap[r] = (u32)old_of.size(); old_of.push_back(r); }
	double *coord = (double *)dalloc((size_t)cap * d_ * sizeof(double));
	u32 *row_slot = (u32 *)dalloc((size_t)cap * 4), *root = (u32 *)dalloc((size_t)cap * 4), *live = (u32 *)dalloc(cap / 8), *ideal = (u32 *)dalloc(cap / 8);
	u32 *inc_off = (u32 *)dalloc((size_t)cap * 4), *inc_len = (u32 *)dalloc((size_t)cap * 4);
	u32 *adj_off = (u32 *)dalloc((size_t)cap * 4), *adj_len = (u32 *)dalloc((size_t)cap * 4);
	u32 *inc_pool = (u32 *)dalloc((size_t)S.cap_inc * 4), *adj_pool = (u32 *)dalloc((size_t)S.cap_adj * 4);
	u32 iu = 0, au = 0;
	for (u32 n = 0; n < old_of.size(); n++) {
		const u32 o = old_of[n];
		for (int j = 0; j < d_; j++) coord[(size_t)j * cap + n] = S.coord[(size_t)j * cap + o];
		row_slot[n] = S.row_slot[o];
		root[n] = S.root[o];
		live[n >> 5] |= 1u << (n & 31);
		if (bit_test(S.ideal, o)) ideal[n >> 5] |= 1u << (n & 31);
		inc_off[n] = iu; inc_len[n] = S.inc_len[o];
		for (u32 q = 0; q < S.inc_len[o]; q++) inc_pool[iu++] = S.inc_pool[S.inc_off[o] + q];
		adj_off[n] = au; adj_len[n] = S.adj_len[o];
		for (u32 q = 0; q < S.adj_len[o]; q++) adj_pool[au++] = remap[S.adj_pool[S.adj_off[o] + q]];
	}
	void *old[] = {S.coord, S.row_slot, S.root, S.live, S.ideal, S.inc_off, S.inc_len, S.adj_off, S.adj_len, S.inc_pool, S.adj_pool};
	for (void *p : old) dfree(p);
	S.coord = coord; S.row_slot = row_slot; S.root = root; S.live = live; S.ideal = ideal; S.inc_off = inc_off; S.inc_len = inc_len;
	S.adj_off = adj_off; S.adj_len = adj_len; S.inc_pool = inc_pool; S.adj_pool = adj_pool;
	memset(S.cls, 0, nrows);
	S.ctl->nrows = S.ctl->n_live = n_live;
	S.ctl->inc_used = iu;
	S.ctl->adj_used = au;
	hdr_ = *S.ctl;
	stats_.compactions++;
}
#endif

void CutEngine::reupload_coords(const double *data_aos, size_t n_slots)
{
	std::vector<u32> slot(hdr_.nrows);
	d2h(slot.data(), S_.row_slot, (size_t)hdr_.nrows * 4);
	std::vector<double> col(hdr_.nrows);
	for (int j = 0; j < d_; j++) {
		d2h(col.data(), S_.coord + (size_t)j * S_.cap_rows, (size_t)hdr_.nrows * sizeof(double));
		for (u32 r = 0; r < hdr_.nrows; r++)
			if (slot[r] < n_slots) col[r] = data_aos[(size_t)slot[r] * d_ + j];
		h2d(S_.coord + (size_t)j * S_.cap_rows, col.data(), (size_t)hdr_.nrows * sizeof(double));
	}
}

void CutEngine::download_structure(HostStructure &o)
{
#ifndef B200_EMULATE
	CK(cudaSetDevice(g_device));
	CK(cudaStreamSynchronize(STREAM));
#endif
	const u32 n = hdr_.nrows;
	o.nrows = n;
	o.row_slot.resize(n); o.inc_off.resize(n); o.inc_len.resize(n); o.adj_off.resize(n); o.adj_len.resize(n);
	o.live_words.resize((n + 31) / 32);
	o.inc_pool.resize(hdr_.inc_used);
	o.adj_pool.resize(hdr_.adj_used);
	d2h(o.row_slot.data(), S_.row_slot, (size_t)n * 4);
	d2h(o.live_words.data(), S_.live, o.live_words.size() * 4);
	d2h(o.inc_off.data(), S_.inc_off, (size_t)n * 4);
	d2h(o.inc_len.data(), S_.inc_len, (size_t)n * 4);
	d2h(o.adj_off.data(), S_.adj_off, (size_t)n * 4);
	d2h(o.adj_len.data(), S_.adj_len, (size_t)n * 4);
	d2h(o.inc_pool.data(), S_.inc_pool, (size_t)hdr_.inc_used * 4);
	d2h(o.adj_pool.data(), S_.adj_pool, (size_t)hdr_.adj_used * 4);
}


// ====================================================================================================
// Wave path: the host side.  The device schedules itself (wave_bodies.h); the host enqueues iterations a few
// ahead, watches a progress record in mapped host memory and steps in only when the scheduler halts: a halfspace
// that has to run alone through the classic path, a capacity that has to grow, a compaction, the end.
// ====================================================================================================
static u32 env_u32(const char *name, u32 dflt)
{
	const char *e = getenv(name);
	return e ? (u32)strtoul(e, nullptr, 10) : dflt;
}
template <class T> static void wave_fresh(T *&p, size_t n)
{
	dfree(p);
	p = (T *)dalloc(n * sizeof(T));
}

// The wave scratch (~250 MB in ~45 arrays) of a killed polytope is parked and adopted by the next polytope of the process
// that starts a device-resident batch on the same device: allocating it costs 45 cudaMalloc calls (12 ms on a good day,
// 100+ ms on a bad one) at the start of the first batch, and freeing it as much at poly__kill.  What is adopted is the
// state the previous owner left, which is what the same engine would find at its next batch: the footprint marks carry
// epochs (monotone, the counter travels with the scratch), everything else is written before it is read -- except the
// K4 column tags, which are facet ids and restart at 0 with a new polytope: those are cleared.  B200_WAVE_PARK=0: off.
#ifdef B200_EMULATE
static const int g_device = 0;
#endif
struct WavePark {
	bool full = false;
	int device = -1;
	WaveDev wd;
	WaveProgress *progress = nullptr;
	u32 rows = 0, epoch = 1;
	u64 rc_cap = 0;
};
static WavePark g_wave_park;
static std::mutex g_wave_park_mu;
static void dzero(void *p, size_t bytes)
{
#ifndef B200_EMULATE
	if (p && bytes) CK(cudaMemset(p, 0, bytes));
#else
	if (p && bytes) memset(p, 0, bytes);
#endif
}
bool CutEngine::wave_adopt()
{
	static const bool on = env_u32_early("B200_WAVE_PARK", 1) != 0;
	if (!on) return false;
	std::lock_guard<std::mutex> lk(g_wave_park_mu);
	WavePark &k = g_wave_park;
	if (!k.full || k.device != g_device || k.wd.cap_he != S_.cap_he || (getenv("B200_WAVE_TRACE") != nullptr) != (k.wd.trace != nullptr)) return false;
	WD_ = k.wd;
	wave_progress_ = k.progress;
	wave_rows_ = k.rows;
	wave_rc_cap_ = k.rc_cap;
	wave_epoch_ = k.epoch;
	k.full = false;
	WD_.nranks = WD_.rank = WD_.shard_min_rows = 0;
	WD_.xsend = WD_.xrecv = WD_.xksend = WD_.xkrecv = nullptr;
	WD_.xflag = WD_.xkflag = nullptr;
	dzero(WD_.facet_epoch, (size_t)B200_WAVE_MAXW * WD_.cap_facets * 4);
	dzero(WD_.fin_ctr, 16);
	return true;
}

void CutEngine::wave_free()
{
	if (WD_.wc && env_u32_early("B200_WAVE_PARK", 1) != 0) {
		std::lock_guard<std::mutex> lk(g_wave_park_mu);
		WavePark &k = g_wave_park;
		if (!k.full) {
			k.full = true;
			k.device = g_device;
			k.wd = WD_;
			k.progress = wave_progress_;
			k.rows = wave_rows_;
			k.rc_cap = wave_rc_cap_;
			k.epoch = wave_epoch_;
			wave_progress_ = nullptr;
			memset(&WD_, 0, sizeof WD_);
			return;
		}
	}
	void *ptrs[] = {WD_.wc, WD_.ctl, WD_.cur, WD_.list, WD_.wflag, WD_.fin_ctr, WD_.trace, WD_.mark, WD_.rc, WD_.vis, WD_.cnt3, WD_.base3, WD_.dead_slots, WD_.he_off, WD_.he_own, WD_.he_inc,
	                WD_.he_k, WD_.he_rank, WD_.he_incpre, WD_.he_flag, WD_.zmask, WD_.padj, WD_.new_padj_off, WD_.new_padj_len, WD_.new_parent, WD_.deg,
	                WD_.adj_fill, WD_.adj_base, WD_.pair_a, WD_.pair_b, WD_.surv_a, WD_.surv_b, WD_.facet_epoch, WD_.facet_local, WD_.dead_facets, WD_.bits};
	for (void *p : ptrs) dfree(p);
#ifndef B200_EMULATE
	if (wave_progress_) cudaFreeHost(wave_progress_);
#else
	free(wave_progress_);
#endif
	wave_progress_ = nullptr;
	memset(&WD_, 0, sizeof WD_);
}

// (re)size the wave scratch; the stream must be idle
void CutEngine::wave_ensure_scratch(u32 n_facets, u32 pairs_per_pos, u64 bits_per_pos)
{
	const size_t L = B200_WAVE_LIST, NP = B200_WAVE_MAXW, NS = B200_WAVE_SLOTS;
	if (!WD_.wc) wave_adopt();
	if (!WD_.wc) {
		WD_.wc = (WaveCtl *)dalloc(sizeof(WaveCtl));
		WD_.ctl = (CutCtl *)dalloc(NS * sizeof(CutCtl));
		WD_.cur = (CutParams *)dalloc(NS * sizeof(CutParams));
		WD_.list = (u32 *)dalloc(NS * L * 4);
		WD_.wflag = (u32 *)dalloc(NS * 4);
		WD_.fin_ctr = (u32 *)dalloc(16);
		if (getenv("B200_WAVE_TRACE")) WD_.trace = (u64 *)dalloc(256 * 64 * 8);
		WD_.cap_he = S_.cap_he;
		WD_.cap_new = S_.cap_he + (u32)L;          // new rows <= half-edges + on-plane copies
		const size_t H = WD_.cap_he, NW = WD_.cap_new;
		WD_.vis = (u32 *)dalloc(NP * L * 4);
		WD_.cnt3 = (u32 *)dalloc(NP * 3 * L * 4);
		WD_.base3 = (u32 *)dalloc(NP * 3 * L * 4);
		WD_.dead_slots = (u32 *)dalloc(NP * L * 4);
		WD_.he_off = (u32 *)dalloc(NP * (L + 1) * 4);
		WD_.he_own = (u32 *)dalloc(NP * H * 4);
		WD_.he_inc = (u32 *)dalloc(NP * H * 4);
		WD_.he_k = (u32 *)dalloc(NP * H * 4);
		WD_.he_rank = (u32 *)dalloc(NP * H * 4);
		WD_.he_incpre = (u32 *)dalloc(NP * H * 4);
		WD_.he_flag = (u8 *)dalloc(NP * H);
		WD_.zmask = (u64 *)dalloc(NP * L * (B200_MAXINC / 64) * 8);
		WD_.padj = (u32 *)dalloc(NP * NW * 4);
		WD_.new_padj_off = (u32 *)dalloc(NP * NW * 4);
		WD_.new_padj_len = (u32 *)dalloc(NP * NW * 4);
		WD_.new_parent = (u32 *)dalloc(NP * NW * 4);
		WD_.deg = (u32 *)dalloc(NP * NW * 4);
		WD_.adj_fill = (u32 *)dalloc(NP * NW * 4);
		WD_.adj_base = (u32 *)dalloc(NP * NW * 4);
#ifndef B200_EMULATE
		CK(cudaHostAlloc((void **)&wave_progress_, sizeof(WaveProgress), cudaHostAllocMapped));
		memset((void *)wave_progress_, 0, sizeof(WaveProgress));
		void *dp = nullptr;
		CK(cudaHostGetDevicePointer(&dp, (void *)wave_progress_, 0));
		WD_.progress = (WaveProgress *)dp;
#else
		wave_progress_ = (WaveProgress *)calloc(1, sizeof(WaveProgress));
		WD_.progress = wave_progress_;
#endif
	}
	if (wave_rows_ < S_.cap_rows) {                // fresh marks read 0, below every tag of a live epoch
		wave_fresh(WD_.mark, S_.cap_rows);
		wave_rows_ = S_.cap_rows;
	}
	if (WD_.cap_facets < n_facets) {
		const u32 cap = (u32)std::max<u64>(n_facets, (u64)WD_.cap_facets * 2);
		wave_fresh(WD_.facet_epoch, NP * cap);    // tags are facet id + 1, never 0
		wave_fresh(WD_.facet_local, NP * cap);
		wave_fresh(WD_.dead_facets, NP * cap);
		WD_.cap_facets = cap;
	}
	if (WD_.cap_pairs < pairs_per_pos) {
		const u32 cap = (u32)std::max<u64>(pairs_per_pos, (u64)WD_.cap_pairs * 2);
		wave_fresh(WD_.pair_a, NP * cap);
		wave_fresh(WD_.pair_b, NP * cap);
		wave_fresh(WD_.surv_a, NP * cap);
		wave_fresh(WD_.surv_b, NP * cap);
		WD_.cap_pairs = cap;
	}
	if (WD_.cap_bits < bits_per_pos) {
		const u64 cap = std::max<u64>(bits_per_pos, WD_.cap_bits * 2);
		wave_fresh(WD_.bits, NP * cap);
		WD_.cap_bits = cap;
	}
}

void CutEngine::wave_sync_ctl(WaveCtl &wc)
{
#ifndef B200_EMULATE
	CK(cudaSetDevice(g_device));
	CK(cudaStreamSynchronize(STREAM));
#endif
	d2h(&wc, WD_.wc, sizeof wc);
	d2h(&hdr_, S_.ctl, sizeof hdr_);
}
void CutEngine::wave_upload_ctl(const WaveCtl &wc)
{
	h2d(WD_.wc, &wc, sizeof wc);
	wave_progress_->halt = wc.halt;
	wave_progress_->iter = wc.iter;
	wave_progress_->done_hs = wc.done_hs;
}

#ifndef B200_EMULATE
template <class K, class... A> static void launch_clusters(K kernel, int n_clusters, int nc, cudaStream_t st, A... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(n_clusters * nc);
	cfg.blockDim = dim3(TAIL_THREADS);
	cfg.stream = st;
	cudaLaunchAttribute at[2];
	int na = 0;
	at[na].id = cudaLaunchAttributeClusterDimension;
	at[na].val.clusterDim.x = nc; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
	na++;
	if (g_pdl) {
		at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[na].val.programmaticStreamSerializationAllowed = 1;
		na++;
	}
	cfg.attrs = at; cfg.numAttrs = na;
	CK(cudaLaunchKernelEx(&cfg, kernel, args...));
}
template <int D> static void launch_wave_classify(const DevState &S, const WaveDev &W, int grid, cudaStream_t st)
{
	launch_dependent(k_wave_classify<D>, grid, K_THREADS, st, S, W);
}

// one iteration (from_stage 0), or only its pair test + adjacency build (1), or only its adjacency build (2)
void CutEngine::wave_enqueue(int from_stage, const double *d_vals, const unsigned char *d_ideal)
{
	const int nclu = wave_max_clusters_, gk4 = num_sms_ * 4;
	if (from_stage == 0) {
		launch_dependent(k_wave_begin, 1, 64, STREAM, S_, WD_, d_vals, d_ideal);
		const int gcl = num_sms_ * 4;
		switch (d_) {
		case 2: launch_wave_classify<2>(S_, WD_, gcl, STREAM); break;
		case 3: launch_wave_classify<3>(S_, WD_, gcl, STREAM); break;
		case 4: launch_wave_classify<4>(S_, WD_, gcl, STREAM); break;
		case 5: launch_wave_classify<5>(S_, WD_, gcl, STREAM); break;
		case 6: launch_wave_classify<6>(S_, WD_, gcl, STREAM); break;
		case 7: launch_wave_classify<7>(S_, WD_, gcl, STREAM); break;
		case 8: launch_wave_classify<8>(S_, WD_, gcl, STREAM); break;
		default: launch_wave_classify<0>(S_, WD_, gcl, STREAM); break;
		}
		if (WD_.xsend) {                    // several GPUs: records of a sharded pass to the peers, theirs into our lists
			launch_dependent(k_wave_xpush, (int)WD_.nranks, K_THREADS, STREAM, S_, WD_);
			launch_dependent(k_wave_xmerge, (int)WD_.nranks * 4, K_THREADS, STREAM, S_, WD_);
			stats_.kernel_launches += 2;
		}
		launch_dependent(k_wave_mark, num_sms_, K_THREADS, STREAM, S_, WD_);
		launch_dependent(k_wave_check, num_sms_, K_THREADS, STREAM, S_, WD_);
		launch_clusters(k_wave_tailA<WAVE_NC>, nclu, WAVE_NC, STREAM, S_, WD_);
		launch_clusters(k_wave_tailB<WAVE_NC>, nclu, WAVE_NC, STREAM, S_, WD_);
		stats_.kernel_launches += 6;
	}
	if (from_stage == 1) { k_wave_k4_reset<<<gk4, K_THREADS, 0, STREAM>>>(S_, WD_); stats_.kernel_launches++; }
	if (from_stage <= 1) {
		launch_dependent(k_wave_k4_filter, gk4, K_THREADS, STREAM, S_, WD_);
		launch_dependent(k_wave_k4_contain, num_sms_ * 8, K_THREADS, STREAM, S_, WD_);
		stats_.kernel_launches += 2;
		if (WD_.xksend) {                   // several GPUs: the adjacent pairs each rank found, to every rank
			launch_dependent(k_wave_k4_xpush, (int)WD_.nranks, K_THREADS, STREAM, S_, WD_);
			launch_dependent(k_wave_k4_xmerge, (int)WD_.nranks * 8, K_THREADS, STREAM, S_, WD_);
			stats_.kernel_launches += 2;
		}
	}
	launch_clusters(k_wave_tail2<WAVE_NC>, nclu, WAVE_NC, STREAM, S_, WD_);
	stats_.kernel_launches++;
	CK(cudaGetLastError());
}
#else
// ---- host-side test double of the wave kernels: the same bodies, run serially
static void emu_wave_classify(const DevState &S, const WaveDev &W)
{
	WaveCtl *w = W.wc;
	if (w->halt) return;
	u32 r_lo = 0, r_hi = w->la_rows;
	if (w->shard) {                      // this rank's row groups, entries into the exchange record
		const u32 ngroups = (w->la_rows + B200_WV_GROUP - 1) / B200_WV_GROUP;
		u32 g_lo, g_hi;
		wave_shard_range(W, ngroups, g_lo, g_hi);
		r_lo = std::min<u32>(w->la_rows, g_lo * B200_WV_GROUP);
		r_hi = std::min<u32>(w->la_rows, g_hi * B200_WV_GROUP);
	}
	for (u32 k = 0; k < w->n_la; k++)
		for (u32 r = r_lo; r < r_hi; r++)
			if (bit_test(S.live, r)) wave_classify_row(S, W, w->la[k], r, w->shard != 0);
	if (!w->shard) return;
	// exchange (all-gather through the host callback), then every record into the lists
	g_comm.callback(W.xsend, W.xrecv, (size_t)B200_X_WORDS * 8);
	for (u32 g = 0; g < W.nranks; g++) {
		const unsigned long long *rec = W.xrecv + (size_t)g * B200_X_WORDS;
		if (rec[1] != w->xseq) fail("multi-rank test double: exchange record of another pass");
		if (rec[0] > B200_X_CAP) { w->halt |= WH_XOVER; continue; }
		for (u32 x = 0; x < (u32)rec[0]; x++) wave_merge_entry(W, rec[2 + x]);
	}
	if (w->halt) wave_publish(W, *w, S.ctl->nrows, S.ctl->n_live);
}
static void emu_wave_form(const DevState &S, const WaveDev &W)
{
	WaveCtl &w = *W.wc;
	if (w.halt) return;
	const u32 nc = wave_candidates(w);
	for (u32 p = 0; p < nc; p++) W.wflag[p] = W.ctl[w.pending[p]].n_list > B200_WAVE_LIST ? 2u : 0u;
	for (u32 p = 0; p < nc; p++)
		for (u32 e = 0, n = std::min<u32>(W.ctl[w.pending[p]].n_list, B200_WAVE_LIST); e < n; e++) wave_mark_entry(S, W, w.pending[p], p, w.epoch, e);
	for (u32 p = 1; p < nc; p++)
		for (u32 e = 0, n = std::min<u32>(W.ctl[w.pending[p]].n_list, B200_WAVE_LIST); e < n; e++) wave_check_entry(S, W, w.pending[p], p, w.epoch, e);
	wave_form_finish(w, W.wflag);
	if (w.halt) wave_publish(W, w, S.ctl->nrows, S.ctl->n_live);
}
static void emu_wave_tailA(const DevState &S0, const WaveDev &W)
{
	const WaveCtl *w = W.wc;
	if (w->halt) return;
	for (u32 q = 0; q < w->n_wave; q++) {
		const DevState S = wave_view(S0, W, w->wave[q], q);
		CutCtl *c = S.ctl;
		const u32 n_list = c->n_list;
		std::vector<std::pair<u32, u8>> vis;
		u32 n_strict = 0;
		for (u32 x = 0; x < n_list; x++) {
			const u32 ent = S.nplist[x], row = ent & B200_WV_ROW_MASK, code = ent >> B200_WV_ROW_BITS;
			if (!bit_test(S.live, row)) continue;
			if ((code & 3u) >= CLS_ZERO) vis.push_back({row, (u8)(code & 3u)});
			n_strict += (code & B200_WV_STRICT) ? 1 : 0;
		}
		std::sort(vis.begin(), vis.end());
		c->status = n_strict ? 0u : (u32)ST_REDUNDANT;
		c->n_strict = n_strict;
		c->min_strict_row = c->min_strict_slot = B200_NONE;
		c->n_zp = c->n_zp_projected = 0;
		c->n_vis = n_strict ? (u32)vis.size() : 0;
		c->n_new = c->inc_new = c->padj_new = 0;
		c->n_minus = c->n_zero = 0;
		c->n_pairs = c->adj_new = c->n_dead_facets = 0;
		c->n_live_scanned = 0;
		c->n_local = c->wl = c->mpad = c->n_surv = 0;
		c->scratch_flag = 0;
		if (!n_strict) { S.facet_alive[S.cur->facet] = 0; continue; }
		u32 H = 0;
		for (u32 i = 0; i < c->n_vis; i++) {
			const u32 r = vis[i].first;
			S.vis[i] = r;
			S.cls[r] = vis[i].second;
			S.he_off[i] = H;
			H += S.adj_len[r];
			for (u32 x = 0; x < (S.inc_len[r] + 63) / 64 && x < B200_MAXINC / 64; x++) S.zmask[(size_t)i * (B200_MAXINC / 64) + x] = 0;
		}
		S.he_off[c->n_vis] = H;
		bool long_zero = false;             // an on-plane vertex on more facets than the mask holds: the cut runs alone
		for (u32 i = 0; i < c->n_vis; i++) long_zero |= S.cls[S.vis[i]] == CLS_ZERO && S.inc_len[S.vis[i]] > B200_MAXINC;
		if (H > S.cap_he || long_zero) { c->status |= ST_NEED_BIG; continue; }
		for (u32 i = 0; i < c->n_vis; i++) he_owner_fill(S, i);
		for (u32 e = 0; e < H; e++) he_eval(S, e);
		u32 carry[3] = {0, 0, 0};
		for (u32 i = 0; i < c->n_vis; i++) {
			he_count(S, i);
			const u8 cl = S.cls[S.vis[i]];
			c->n_minus += (cl == CLS_MINUS);
			c->n_zero += (cl == CLS_ZERO);
			for (int k = 0; k < 3; k++) { S.base3[3 * (size_t)i + k] = carry[k]; carry[k] += S.cnt3[3 * (size_t)i + k]; }
		}
		c->n_new = carry[0]; c->inc_new = carry[1]; c->padj_new = carry[2];
		if (carry[0] > W.cap_new || carry[2] > W.cap_new) c->status |= ST_OVF_PADJ;
	}
}
static void emu_wave_tailB(const DevState &S0, const WaveDev &W)
{
	WaveCtl &w = *W.wc;
	if (w.halt || w.n_wave == 0) return;
	WaveCut cut[B200_WAVE_MAXW];
	for (u32 q = 0; q < w.n_wave; q++) wave_gather_cut(W, w, q, cut[q]);
	WavePlan pl;
	wave_plan(w, cut, S0.ctl->nrows, S0.ctl->inc_used, S0.ctl->n_live, S0.cap_rows, S0.cap_inc, W.cap_bits, pl);
	std::vector<u32> tgt;
	for (u32 p = 0; p < w.n_pending; p++) {
		bool done = false;
		for (u32 q2 = 0; q2 < pl.n_commit; q2++) done |= (w.wave[q2] == w.pending[p]);
		if (!done) tgt.push_back(w.pending[p]);
	}
	const u32 slot_cnt0 = S0.ctl->slot_cnt, nrows0 = S0.ctl->nrows;
	for (u32 q = w.n_wave; q-- > 0;) {            // any order is valid: run the wave backwards here
		const DevState S = wave_view(S0, W, w.wave[q], q);
		CutCtl *c = S.ctl;
		const bool redundant = (c->status & ST_REDUNDANT) != 0;
		if (q >= pl.n_commit) {
			if (!redundant) for (u32 i = 0; i < c->n_vis; i++) reset_class(S, i);
			c->status |= ST_WAVE_DEFER;
			if (q == 0) {
				w.n_commit = 0;
				w.halt |= pl.halt;
				w.halt_hs = pl.halt_hs;
				w.halt_rows = pl.need_rows;
				w.halt_inc = pl.need_inc;
				w.halt_bits = pl.need_bits;
				wave_publish(W, w, S0.ctl->nrows, S0.ctl->n_live);
			}
			continue;
		}
		if (q == 0) w.n_commit = pl.n_commit;
		c->n_live = pl.live_before[q];
		if (redundant) continue;
		c->nrows = pl.rows_base[q];
		c->slot_cnt = slot_cnt0 + (pl.rows_base[q] - nrows0);
		c->inc_used = pl.inc_base[q];
		S.facet_cnt[S.cur->facet] = c->n_new;
		const CutParams &P = *S.cur;
		const u32 H = S.he_off[c->n_vis];
		for (u32 e = H; e-- > 0;) he_emit(S, P, e);
		for (u32 i = 0; i < c->n_vis; i++) he_finish_vertex(S, P, i);
		for (u32 i = 0; i < c->n_vis; i++) collect_dead_facets(S, i);
		for (u32 t : tgt)
			for (u32 j = 0; j < c->n_new; j++) wave_classify_row(S, W, t, c->nrows + j);
		k4_plan(S);
		for (u64 x = 0; x < (u64)c->n_local * (c->mpad / 64); x++) k4_zero_cols(S, x);
		for (u32 j = 0; j < c->n_new; j++) k4_build_row(S, j);
	}
}
static void emu_wave_k4(const DevState &S0, const WaveDev &W, bool reset)
{
	WaveCtl *w = W.wc;
	if (w->halt) return;
	const bool shk = w->shard_k4 != 0;
	if (shk && reset) wave_k4_record_reset(W, w->xseq_k);
	for (u32 q = 0; q < w->n_commit; q++) {
		const DevState S = wave_view(S0, W, w->wave[q], q);
		CutCtl *c = S.ctl;
		if (c->status & (ST_REDUNDANT | ST_OVF_A | ST_OVF_B | ST_ERR_DEGENERATE | ST_NEED_BIG)) continue;
		if (reset) { c->n_surv = c->n_pairs = 0; for (u32 j = 0; j < c->n_new; j++) S.deg[j] = 0; }
		const u32 M = c->n_new;
		for (u32 a = 0; a < M; a++)
			for (u32 b = a + 1; b < M; b++)
				if (!shk || (a + b) % W.nranks == W.rank) k4_filter_pair(S, a, b);      // (sharded: this rank's share of the pair space)
		if (c->n_surv > S.cap_pairs) {
			if (shk) { W.xksend[2] |= 1u; if (W.xksend[3] < c->n_surv) W.xksend[3] = c->n_surv; }
			continue;
		}
		for (u32 sv = 0; sv < c->n_surv; sv++) {
			const u32 a = S.surv_a[sv], b = S.surv_b[sv];
			if (!k4_adjacent_by_columns(S, a, b, c->n_new, c->wl, c->mpad)) continue;
			if (shk) wave_k4_send_pair(W, q, a, b);
			else { wave_flag_adjacent(S, sv, a, b); c->n_pairs++; }
		}
	}
	if (!shk) return;
	// exchange of the adjacent pairs (all-gather through the host callback), then every pair under its cut
	g_comm.callback(W.xksend, W.xkrecv, (size_t)B200_XK_WORDS * 8);
	u32 need = 0;
	bool over_rec = false, over_surv = false;
	for (u32 g = 0; g < W.nranks; g++) {
		const unsigned long long *rec = W.xkrecv + (size_t)g * B200_XK_WORDS;
		if (rec[1] != w->xseq_k) fail("multi-rank test double: pair record of another exchange");
		if ((u32)rec[0] > B200_XK_CAP) over_rec = true;
		if (rec[2] & 1u) { over_surv = true; need = std::max<u32>(need, (u32)rec[3]); }
	}
	if (over_surv) { w->halt |= WH_GROW_PAIRS; w->halt_pairs = std::max(w->halt_pairs, need); }
	else if (over_rec) w->halt |= WH_XOVER_K4;
	if (w->halt) { wave_publish(W, *w, S0.ctl->nrows, S0.ctl->n_live); return; }
	for (u32 g = 0; g < W.nranks; g++) {
		const unsigned long long *rec = W.xkrecv + (size_t)g * B200_XK_WORDS;
		for (u32 x = 0; x < (u32)rec[0]; x++) wave_k4_merge_pair(S0, W, *w, rec[4 + x]);
	}
}
static void emu_wave_tail2(const DevState &S0, const WaveDev &W)
{
	WaveCtl &w = *W.wc;
	if (w.halt || w.n_commit == 0) return;
	WaveCut cut[B200_WAVE_MAXW];
	u32 adj_base[B200_WAVE_MAXW], adj_new[B200_WAVE_MAXW], need_adj, need_pairs;
	for (u32 q = 0; q < w.n_commit; q++) wave_gather_cut(W, w, q, cut[q]);
	const u32 fl = wave_adj_plan(cut, w.n_commit, S0.ctl->adj_used, S0.cap_adj, W.cap_pairs, adj_base, adj_new, need_adj, need_pairs);
	for (u32 q = 0; q < w.n_commit; q++) { W.ctl[w.wave[q]].adj_used = adj_base[q]; W.ctl[w.wave[q]].adj_new = adj_new[q]; }
	if (fl & 8u) { w.halt |= WH_GROW_PAIRS; w.halt_pairs = need_pairs; wave_publish(W, w, S0.ctl->nrows, S0.ctl->n_live); return; }
	if (fl & 16u) { w.halt |= WH_GROW_ADJ; w.halt_adj = need_adj; wave_publish(W, w, S0.ctl->nrows, S0.ctl->n_live); return; }
	for (u32 q = 0; q < w.n_commit; q++) {
		const DevState S = wave_view(S0, W, w.wave[q], q);
		CutCtl *c = S.ctl;
		if (c->status & ST_REDUNDANT) continue;
		u32 carry = 0;
		for (u32 j = 0; j < c->n_new; j++) { S.adj_base[j] = carry; carry += S.new_padj_len[j] + S.deg[j]; }
		if (carry != c->adj_new) fail("wave adjacency plan disagrees with the scan");
		for (u32 j = 0; j < c->n_new; j++) adj_place(S, j);
		if (w.shard_k4) for (u32 pp = 0; pp < c->n_pairs; pp++) adj_pair_fill(S, pp);
		else for (u32 sv = 0; sv < c->n_surv; sv++) adj_pair_fill_surv(S, sv);
		for (u32 j = 0; j < c->n_new; j++) adj_sort(S, j);
	}
}
static void emu_wave_begin(const DevState &S, const WaveDev &W, const double *vals, const unsigned char *ideal)
{
	WaveCtl &w = *W.wc;
	if (w.halt) return;
	const u32 n_commit = w.n_commit;
	if (n_commit) {
		WaveCut cut[B200_WAVE_MAXW];
		int rc[B200_WAVE_MAXW];
		for (u32 q = 0; q < n_commit; q++) wave_gather_cut(W, w, q, cut[q]);
		CutCtl m = *S.ctl;
		wave_commit(w, m, cut, S.d, rc);
		*S.ctl = m;
		for (u32 q = 0; q < n_commit; q++) W.rc[cut[q].hs] = rc[q];
	}
	w.iter++;
	if (!w.halt) {
		wave_la_plan(w, S.ctl->nrows);
		wave_shard_plan(w, W, S.ctl->nrows);
		if (w.shard) { W.xsend[0] = 0; W.xsend[1] = w.xseq; }
		if (w.shard_k4) { wave_k4_record_reset(W, ++w.xseq_k); w.st_sharded_k4++; }
		for (u32 k = 0; k < w.n_la; k++) wave_la_init(S, W, w, k, vals, ideal);
	}
	wave_publish(W, w, S.ctl->nrows, S.ctl->n_live);
}
void CutEngine::wave_enqueue(int from_stage, const double *d_vals, const unsigned char *d_ideal)
{
	if (from_stage == 0) {
		emu_wave_begin(S_, WD_, d_vals, d_ideal);
		emu_wave_classify(S_, WD_);
		emu_wave_form(S_, WD_);
		emu_wave_tailA(S_, WD_);
		emu_wave_tailB(S_, WD_);
	}
	if (from_stage <= 1) emu_wave_k4(S_, WD_, from_stage == 1);
	emu_wave_tail2(S_, WD_);
}
#endif

long CutEngine::cut_batch_from_device(const double *d_vals, const unsigned char *d_ideal, u64 n, u32 facet0, u32 batch_first, int *rc_out)
{
	static const u32 min_live = env_u32("B200_WAVE_MIN_LIVE", 20000);    // below: one launch per cut (classic path) wins
	const bool tiny_test = (flags_ & 32) != 0;                            // test hook: waves from the first halfspace on
	const bool enabled = (env_u32("B200_WAVES", 1) != 0 || tiny_test) && !(flags_ & (4 | 64)) && n < B200_WV_ROW_MASK;
	std::vector<int> rc_host(n, -1);
	long cuts = 0;
	u64 i = 0;
	auto serial = [&](u64 hs) {
		const int rc = cut_from_device(d_vals, d_ideal, hs, facet0 + (u32)hs, batch_first);
		rc_host[hs] = rc;
		cuts += (rc == 0);
	};
	while (i < n && (!enabled || (!tiny_test && hdr_.n_live < min_live))) serial(i++);
	if (i < n) {
		// ---- wave mode from halfspace i on
#ifndef B200_EMULATE
		CK(cudaSetDevice(g_device));
		CK(cudaStreamSynchronize(STREAM));
		if (!wave_max_clusters_) {
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(B200_WAVE_MAXW * WAVE_NC);
			cfg.blockDim = dim3(TAIL_THREADS);
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension;
			at[0].val.clusterDim.x = WAVE_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
			cfg.attrs = at; cfg.numAttrs = 1;
			int na = 0, nb = 0, nc2 = 0;
			CK(cudaOccupancyMaxActiveClusters(&na, k_wave_tailA<WAVE_NC>, &cfg));
			CK(cudaOccupancyMaxActiveClusters(&nb, k_wave_tailB<WAVE_NC>, &cfg));
			CK(cudaOccupancyMaxActiveClusters(&nc2, k_wave_tail2<WAVE_NC>, &cfg));
			wave_max_clusters_ = std::max(1, std::min(std::min(na, nb), std::min(nc2, (int)B200_WAVE_MAXW)));
		}
#else
		wave_max_clusters_ = B200_WAVE_MAXW;
#endif
		ensure_facets(facet0 + (u32)n + 1);
		{
			// pair-test scratch of one wave position: bit matrices for a cut of up to 2048 new rows with every facet a
			// column (grown on demand beyond that), 2^17 candidate pairs
			const u32 wl_ub = (facet0 + (u32)n + 64) / 64;
			wave_ensure_scratch(facet0 + (u32)n + 1, std::max<u32>(1u << 17, WD_.cap_pairs), std::max<u64>(k4_words(wl_ub, 2048, wl_ub * 64), WD_.cap_bits));
		}
		if (wave_rc_cap_ < n) { wave_fresh(WD_.rc, n); wave_rc_cap_ = n; }
		// several ranks: the look-ahead passes over large polytopes are sharded, exchanged over peer-mapped memory
		WD_.nranks = (u32)nranks_;
		WD_.rank = (u32)rank_;
		WD_.shard_min_rows = tiny_test ? 0u : env_u32("B200_SHARD_MIN_ROWS_WAVE", 3000000);
		if (nranks_ > 1) {
#ifndef B200_EMULATE
			xarea_setup(STREAM);
#else
			xarea_setup();
#endif
			if (g_x.ok) {
				WD_.xsend = g_x.send; WD_.xrecv = g_x.recv; WD_.xflag = g_x.flag;
				WD_.xksend = g_x.ksend; WD_.xkrecv = g_x.krecv; WD_.xkflag = g_x.kflag;
				for (int g = 0; g < B200_X_MAXRANKS; g++) {
					WD_.xpeer_recv[g] = g_x.peer_recv[g]; WD_.xpeer_flag[g] = g_x.peer_flag[g];
					WD_.xkpeer_recv[g] = g_x.kpeer_recv[g]; WD_.xkpeer_flag[g] = g_x.kpeer_flag[g];
				}
			}
		}
		WaveCtl wc;
		memset(&wc, 0, sizeof wc);
		wc.xseq = g_x.seq;
		wc.xseq_k = g_x.seq_k;
		// (off by default: at the cut sizes of the measured workloads -- ~10^3 new vertices per cut -- the pair test of a
		// wave is 40 us of latency-bound work and one more cross-rank rendezvous per iteration costs more than 7/8 of it
		// saves: 46.4 k vs 48.5 k cuts/s on 8 ranks, 48.3 k vs 52.1 k on 2; it is for cuts with 10^4 new vertices)
		wc.shard_k4 = (WD_.xksend && env_u32("B200_K4_SHARD", 0) != 0) ? 1u : 0u;
		wc.n_total = (u32)n;
		wc.facet0 = facet0;
		wc.batch_first = batch_first;
		wc.max_wave = std::max<u32>(1, std::min<u32>(env_u32("B200_WAVE_MAX", B200_WAVE_MAXW), (u32)wave_max_clusters_));
		wc.cand = std::max<u32>(1, std::min<u32>(env_u32("B200_WAVE_CAND", 24), B200_WAVE_SLOTS));
		wc.in_order = env_u32("B200_WAVE_IN_ORDER", 0);
		wc.refill_below = std::max<u32>(1, std::min<u32>(env_u32("B200_WAVE_REFILL", 20), B200_WAVE_SLOTS));
		wc.next_hs = wc.done_hs = (u32)i;
		wc.epoch = wave_epoch_;
		for (u32 s = 0; s < B200_WAVE_SLOTS; s++) wc.slot_hs[s] = B200_NONE;
		wave_upload_ctl(wc);
		const u32 depth = std::max<u32>(1, env_u32("B200_WAVE_DEPTH", 3));
		u32 issued = 0;                       // iterations enqueued so far (compared with the device's count of completed ones)
		int stage = 0;
		u64 spins = 0, halts = 0;
		for (;;) {
			const u32 it = wave_progress_->iter, hl = wave_progress_->halt;
			if (!hl) {
				if (issued - it < depth) {
					const double te = now_us();
					wave_enqueue(stage, d_vals, d_ideal);
					stats_.host_us[7] += now_us() - te;
					stage = 0;
					issued++;
				}
#ifndef B200_EMULATE
				else if ((++spins & 0xffff) == 0) {         // make sure the stream is still healthy
					cudaError_t e = cudaStreamQuery(STREAM);
					if (e != cudaSuccess && e != cudaErrorNotReady) fail(std::string("CUDA error while a wave was running: ") + cudaGetErrorString(e));
					if (e == cudaSuccess && wave_progress_->iter == it && !wave_progress_->halt && issued - it >= depth)
						fail("bensolve_b200: the wave scheduler stopped without a halt record");
				}
#endif
				continue;
			}
			// ---- the scheduler halted: everything enqueued behind the halt returns at once
			wave_sync_ctl(wc);
			stats_.wave_halts++;
			issued = wc.iter;
			stage = 0;
			if (wc.halt & WH_DONE) break;
			if (++halts > 8 * n + 64) fail("bensolve_b200: the wave scheduler does not make progress");
			if (wc.halt & WH_SERIAL) {
				if (wc.n_pending == 0 || wc.slot_hs[wc.pending[0]] != wc.halt_hs) fail("bensolve_b200: wave scheduler state is inconsistent");
				const u32 slot = wc.pending[0];
				serial(wc.halt_hs);                                   // (may grow and compact)
				stats_.wave_serial++;
				wc.slot_hs[slot] = B200_NONE;
				for (u32 p = 1; p < wc.n_pending; p++) wc.pending[p - 1] = wc.pending[p];
				wc.n_pending--;
				wc.done_hs++;
				wc.reclassify = 1;
			} else if (wc.halt & WH_GROW) {
				if (wc.halt_rows > S_.cap_rows) ensure_rows((u32)std::min<u64>(0xFFFF0000ull, std::max<u64>((u64)wc.halt_rows + 4096, (u64)S_.cap_rows + S_.cap_rows / 2)));
				if (wc.halt_inc > S_.cap_inc) ensure_inc((u32)std::min<u64>(0xFFFF0000ull, std::max<u64>(wc.halt_inc, (u64)S_.cap_inc + S_.cap_inc / 2)));
				if (wc.halt_bits > WD_.cap_bits) wave_ensure_scratch(WD_.cap_facets, WD_.cap_pairs, wc.halt_bits);
			} else if (wc.halt & WH_GROW_ADJ) {
				ensure_adj(wc.halt_adj);
				stage = 2;
			} else if (wc.halt & WH_GROW_PAIRS) {
				wave_ensure_scratch(WD_.cap_facets, wc.halt_pairs + wc.halt_pairs / 2, WD_.cap_bits);
				stage = 1;
				if (wc.shard_k4) wc.xseq_k++;           // the redone pair test is a new exchange
			} else if (wc.halt & WH_XOVER_K4) {         // (on every rank alike) this batch goes on with the pair test replicated
				wc.shard_k4 = 0;
				stage = 1;
			} else if (wc.halt & WH_COMPACT) {
				compact();
				wc.reclassify = 1;
			} else if (wc.halt & WH_XFAIL) {
				fail("bensolve_b200: a peer rank's look-ahead record did not arrive (rank died or lost its device)");
			} else if (wc.halt & WH_XOVER) {       // (on every rank alike) rebuild the lists of this pass unsharded
				wc.noshard_once = 1;
				wc.reclassify = 1;
			}
			if (wc.done_hs >= wc.n_total) break;
			wave_ensure_scratch(WD_.cap_facets, WD_.cap_pairs, WD_.cap_bits);   // the mark array follows the row capacity
			wc.halt = 0;
			wave_upload_ctl(wc);
		}
		wave_epoch_ = wc.epoch + 1;
		g_x.seq = wc.xseq;
		g_x.seq_k = wc.xseq_k;
		// results and statistics of the wave run
		std::vector<int> rc_dev(n);
		d2h(rc_dev.data(), WD_.rc, n * sizeof(int));
		for (u64 hs = i; hs < n; hs++)
			if (rc_host[hs] < 0) rc_host[hs] = rc_dev[hs];
		cuts += (long)wc.st_cuts;
		stats_.cuts += wc.st_cuts; stats_.redundant += wc.st_redundant; stats_.vertex_evals += wc.st_evals; stats_.rows_scanned += wc.st_rows_scanned;
		stats_.minus += wc.st_minus; stats_.zero += wc.st_zero; stats_.edge_vertices += wc.st_edge; stats_.copies += wc.st_copies;
		stats_.pair_tests += wc.st_pair_tests; stats_.new_adjacent_pairs += wc.st_pairs; stats_.algorithmic_bytes += wc.st_bytes;
		stats_.waves += wc.st_waves; stats_.wave_cuts += wc.st_cuts + wc.st_redundant; stats_.la_passes += wc.st_la_passes; stats_.wave_deferred += wc.st_deferred; stats_.sharded_passes += wc.st_sharded; stats_.sharded_pair_tests += wc.st_sharded_k4;
		small_dirty_ = true;
		expect_vis_ = expect_m_ = 0;
		if (WD_.trace) {            // start of each kernel of the last iterations, relative to the iteration's first kernel
			std::vector<u64> tr(256 * 64);
			d2h(tr.data(), WD_.trace, tr.size() * 8);
			const u32 last = wc.iter;
			u64 prev_first = 0;
			for (u32 it = last > 24 ? last - 24 : 0; it < last; it++) {
				const u64 *t = tr.data() + ((it & 255u) << 6);
				fprintf(stderr, "[b200] trace iter %u: +%.1f | kernels", it, prev_first ? (double)(t[0] - prev_first) / 1e3 : 0.0);
				for (int k = 1; k < 9; k++) fprintf(stderr, " %.1f", (double)((long long)(t[k] - t[0])) / 1e3);
				const int grp[4][2] = {{16, 22}, {24, 29}, {32, 36}, {40, 45}};
				const int kern[4] = {4, 5, 8, 0};
				const char *nm[4] = {"tailA", "tailB", "tail2", "begin"};
				for (int g = 0; g < 4; g++) {
					fprintf(stderr, " | %s", nm[g]);
					for (int k = grp[g][0]; k <= grp[g][1]; k++) fprintf(stderr, " %.1f", (double)((long long)(t[k] - t[kern[g]])) / 1e3);
				}
				fprintf(stderr, "\n");
				prev_first = t[0];
			}
		}
		if (getenv("B200_PHASES"))
			fprintf(stderr, "[b200] waves: %llu iterations, %llu cuts (%.2f per wave), %llu look-ahead passes, %llu deferred, %llu serial, %llu halts; device %.1f us per iteration (first to last commit), host enqueue %.1f us per iteration\n",
			        (unsigned long long)wc.st_waves, (unsigned long long)(wc.st_cuts + wc.st_redundant), (double)(wc.st_cuts + wc.st_redundant) / std::max<u64>(1, wc.st_waves),
			        (unsigned long long)wc.st_la_passes, (unsigned long long)wc.st_deferred, (unsigned long long)stats_.wave_serial, (unsigned long long)stats_.wave_halts,
			        (double)(wc.t_last - wc.t_first) / 1e3 / std::max<u64>(1, wc.st_waves), stats_.host_us[7] / std::max<u64>(1, wc.st_waves));
	}
	if (rc_out) for (u64 hs = 0; hs < n; hs++) rc_out[hs] = rc_host[hs];
	return cuts;
}
