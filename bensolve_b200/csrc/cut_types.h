// Device-side data layout of the cut engine (see DESIGN.md "Data layout in HBM").
//
// One CutEngine instance mirrors one reference `poly_args` (bslv_poly.h:71-82).  The primal
// polytope lives in HBM as
//   * coord    : SoA FP64, coord[j*cap_rows + r]   (coalesced along the vertex axis for K1)
//   * live/ideal : 32-bit-word bitsets over device rows
//   * row_slot : device row -> host slot number (rows are compacted, host slots never are)
//   * inc_pool : per-row SORTED facet-id lists (immutable after creation; bslv_poly.c keeps them
//                as unsorted malloc'd size_t lists, bslv_poly.h:49-53)
//   * adj_pool : per-row neighbour rows (entries are rewritten in place when a neighbour is cut off)
//   * facet_cnt: live vertices per facet (replaces dual.incidence[f].cnt tests, bslv_poly.c:686,698)
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

#define B200_MAXD 16        // highest supported dimension q of the image space
#define B200_MAXINC 1024    // longest incidence list an on-plane (ZERO) vertex may have (degenerate inputs)
#define B200_NONE 0xFFFFFFFFu

#if defined(__CUDACC__) && !defined(B200_EMULATE)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

// vertex classes with respect to the current halfspace
enum : u8 { CLS_PLUS = 0, CLS_ZP = 1, CLS_ZERO = 2, CLS_MINUS = 3, CLS_DEAD = 4 };

// status bits of a cut
enum : u32 {
	ST_REDUNDANT = 1u,      // no vertex strictly violates: poly__add_vrtx returns EXIT_FAILURE
	ST_OVF_ROWS = 2u,       // row capacity too small          (detected before any mutation)
	ST_OVF_INC = 4u,        // incidence pool too small        (idem)
	ST_OVF_PADJ = 8u,       // PLUS-neighbour scratch too small (idem)
	ST_OVF_PAIRS = 16u,     // K4 pair buffer too small        (K4 is re-runnable)
	ST_OVF_ADJ = 32u,       // adjacency pool too small        (adjacency build is re-runnable)
	ST_ERR_DEGENERATE = 64u,// an incidence list exceeds B200_MAXINC
	ST_OVF_BITS = 128u,     // K4 bit matrix too small         (K4 is re-runnable)
	ST_OVF_STAGE = 256u,    // delta staging buffer too small  (packing is re-runnable)
	ST_NEED_BIG = 512u,     // cut too large for the single-CTA tail: rerun through the multi-kernel path (nothing mutated)
	ST_K4_PENDING = 1024u   // tail stopped before K4's pair test (too many new vertices for one CTA): run k4_filter/k4_contain/k_tail2
};
#define ST_OVF_A (ST_OVF_ROWS | ST_OVF_INC | ST_OVF_PADJ)
#define ST_OVF_B (ST_OVF_PAIRS | ST_OVF_ADJ | ST_OVF_BITS)

// The halfspace of the current cut, h.x >= alpha (alpha replaced by 0 for ideal vertices), with
// the reference's three thresholds pre-added in FP64 exactly as bslv_poly.c:126,573,596,666 does.
struct CutParams {
	double h[B200_MAXD];
	double alpha;
	double hi[2];   // thr + 1e-9        index 1 = ideal vertex (thr = 0)
	double mid[2];  // thr + 1e-2*1e-9
	double lo[2];   // thr - 1e-9
	double hh;      // sum h_j^2, left to right (bslv_poly.c:668-670)
	u32 facet;      // id of the new facet = dual slot of this halfspace
	u32 batch_first; // primal.cnt when the host mirror was last coherent (sltn inheritance roots, bslv_poly.c:583-587)
	u32 seq;         // sequence number the device publishes in the staged header when the record is complete
	u32 zp_done;     // 1: rerun of a cut that bailed out (capacity, size) AFTER its ZERO+ closure had projected rows in place
	                 // (bslv_poly.c:666-674): the closure activates the same rows again but must not project them twice
	double h1;       // sum |h_j| (wave path only: scale of the guard band of the look-ahead classification)
};

// Counters shared by the kernels of one cut; the host reads it back once per cut.
struct CutCtl {
	// persistent
	u32 nrows;        // device rows in use (live + dead, before this cut's appends)
	u32 slot_cnt;     // host-visible primal.cnt
	u32 n_live;       // live vertices
	u32 inc_used, adj_used;
	// per cut
	u32 status;
	u32 n_strict;       // vertices with h.x < thr - eps (trigger, bslv_poly.c:126)
	u32 min_strict_row; // lowest such row (its slot is what the reference leaves in args->idx)
	u32 n_zp;           // ZERO+ candidates seen by K1
	u32 n_zp_projected;
	u32 n_vis;          // non-PLUS rows (compact list length)
	u32 n_new, inc_new, padj_new;
	u32 n_minus, n_zero;
	u32 n_pairs;        // adjacent pairs found by K4
	u32 adj_new;
	u32 n_dead_facets;
	u32 n_live_scanned; // live rows classified by K1
	u32 min_strict_slot;
	// K4 (bitset form)
	u32 n_local;        // distinct facets (other than the new one) met on the new facet's vertices
	u32 wl;             // 64-bit words per bit-matrix row = ceil(n_local/64)
	u32 mpad;           // n_new rounded up to 32: column stride of the word-major bit matrix
	u32 n_surv;         // pairs that pass the AND+POPC filter
	u32 stage_bytes;    // size of the packed delta record (header included)
	u32 scratch_flag;   // cluster-wide 'something changed' flag of the ZERO+ closure
	u32 n_list;         // non-PLUS rows K1 appended to `nplist` (unordered; > B200_VIS_MAX = overflow)
	u32 reserved0;
};
#define B200_STAGE_HDR 256u   // the packed delta starts with two copies of CutCtl, 128 bytes each: the final header (status of the whole cut)
#define B200_STAGE_EARLY 128u // and an early one, published when the payload behind it is complete but the adjacency build still runs
#define B200_STAGE_SEQ 124u   // byte offset of the 'record complete' sequence number inside the header


struct DevState {
	int d;
	u32 cap_rows, cap_inc, cap_adj, cap_facets, cap_padj, cap_pairs, cap_tiles;
	u32 cap_he;          // half-edges the tail kernels' scratch holds (<= B200_HE_CAP; env B200_HE_CAP lowers it, test hook)
	u64 cap_bits;
	double *coord;
	u32 *row_slot, *live, *ideal;
	u32 *root;           // per row: slot (< batch_first) whose sltn flag / pre-image this row inherits, or NONE
	u32 *inc_off, *inc_len, *adj_off, *adj_len;
	u32 *inc_pool, *adj_pool;
	u32 *facet_cnt, *facet_alive;
	u8 *cls;
	// per-cut scratch
	u32 *tile_cnt, *tile_base;
	u32 *vis;            // [cap_rows] non-PLUS rows, ascending
	u32 *cnt3, *base3;   // [3*cap_rows] (new rows, incidence entries, PLUS neighbours) per visited row
	u32 *padj;           // [cap_padj] PLUS neighbours of the new rows
	u32 *new_padj_off, *new_padj_len, *new_parent, *deg, *adj_fill, *adj_base; // [cap_rows], by new-row index
	u32 *pair_a, *pair_b; // [cap_pairs] adjacent pairs (new-row indices)
	u32 *surv_a, *surv_b; // [cap_pairs] pairs that passed the popcount filter
	u32 *facet_epoch, *facet_local; // [cap_facets] K4 column relabelling, valid where epoch == facet+1
	u64 *bits;            // [cap_bits] K4 incidence bit matrix, word-major: bits[w*mpad + new_row]
	u32 *dead_slots;     // [cap_rows] by visited index
	// tail-kernel scratch (small cuts): per-tile lists of non-PLUS rows written by the streaming K1,
	// and the half-edge work items (visited vertex, adjacency slot)
	u32 *nplist;         // [B200_VIS_MAX] non-PLUS rows in the order K1's atomics produced; the tail rank-sorts them into `vis`
	u32 *he_off;         // [B200_VIS_MAX + 1]
	u32 *he_own, *he_inc; // [B200_HE_CAP]
	u32 *he_k, *he_rank, *he_incpre; // [B200_HE_CAP] neighbour row; rank / incidence offset among the vertex's PLUS half-edges
	u8 *he_flag;         // [B200_HE_CAP]
	u64 *zmask;          // [B200_VIS_MAX * B200_MAXINC/64] shared-facet masks of ZERO vertices
	u64 *zlong;          // [B200_VIS_MAX * zlong_words] (or null) facet bitmaps of ZERO vertices whose lists are longer than the mask:
	                     // bit f of row i <=> facet f lies on a PLUS neighbour of visited entry i.  All zero between cuts.
	u32 zlong_words;     // words per row = facets the bitmap covers / 64
	u32 *dead_facets;    // [cap_facets]
	unsigned char *stage; // [cap_stage] packed per-cut delta: header | coords AoS | parent | ideal | dead slots | dead facets.
	                      // Device alias of MAPPED PINNED HOST memory: the kernels write the record straight to the host.
	u64 cap_stage;
	CutCtl *ctl;
	CutParams *cur;
	// multi-GPU exchange (state replicated, K1 sharded by row range): per-rank record = XchgHeader + entries
	u32 *xchg_send;      // [B200_XCHG_WORDS]
	u32 *xchg_recv;      // [nranks * B200_XCHG_WORDS]
	u64 *dbg;            // [32] phase time stamps of the tail kernel (diagnostics, flags bit0)
};

#define B200_TILE 2048u  // rows per K1/K2 tile
#define B200_VIS_MAX 4096u // most visited vertices the single-CTA tail handles
#define B200_HE_CAP 65536u // most half-edges the single-CTA tail handles
#define B200_K4_SMALL 256u
#define B200_XCHG_CAP 4096u   // most non-PLUS rows one rank can report per cut (else the multi-kernel path runs unsharded)
#define B200_XCHG_WORDS (4u + B200_XCHG_CAP)  // header {n_strict, min_strict_row, n_zp, n_entries} + entries (row | class << 30) // most new vertices whose pair test the single-CTA tail does itself


// ====================================================================================================
// Wave path (device-resident batches): look-ahead classification + concurrent independent cuts.
//
// The halfspaces of a batch are all known, so (1) ONE pass over the coordinates classifies every row against
// up to B200_WAVE_SLOTS pending halfspaces (look-ahead K1) and keeps, per halfspace, the short list of rows that
// are not safely PLUS; each later cut classifies only the few hundred rows it creates against the halfspaces
// still pending; (2) cuts whose neighbourhoods do not touch commute, so a *wave* of such cuts runs concurrently,
// one thread-block cluster per cut, through the same phases as a single cut.  DESIGN.md section 4 has the proof
// sketch of the commutation rule (guard band, footprint marks).
// ====================================================================================================
#define B200_WAVE_SLOTS 32u      // pending halfspaces that hold a look-ahead list
#define B200_WAVE_MAXW 16u       // most cuts in one wave (also bounded by the co-resident clusters of the device)
#define B200_WAVE_LIST B200_VIS_MAX   // entries per look-ahead list (longer: that cut runs alone through the classic path)
// list entry = row | code << 28; code bits 0-1 = class (CLS_PLUS here means: PLUS, but inside the guard band -- listed
// for the conflict test only), bit 2 = strictly violated (t < thr - eps, the trigger of bslv_poly.c:126)
#define B200_WV_ROW_BITS 28
#define B200_WV_ROW_MASK 0x0FFFFFFFu
#define B200_WV_STRICT 4u
#define B200_WV_GUARD 1e-7       // relative width of the guard band (rounding errors of a new vertex are ~1e-15 relative)

enum : u32 {                     // WaveCtl::halt
	WH_DONE = 1u,                // every halfspace of the batch is processed
	WH_SERIAL = 2u,              // halfspace halt_hs must run alone through the classic path (ZERO+ rows, list / half-edge / scratch overflow)
	WH_GROW = 4u,                // rows / incidence pool too small for the next wave: halt_rows / halt_inc say how much is needed
	WH_GROW_ADJ = 8u,            // adjacency pool too small: grow, then redo only the adjacency build of this wave
	WH_COMPACT = 16u,            // dead rows outnumber live ones: compact, rebuild the look-ahead lists
	WH_GROW_PAIRS = 32u,         // pair buffers of a wave position too small: grow, then redo the pair test and the adjacency build
	WH_XOVER = 64u,              // a rank's exchange record overflowed: the lists of this pass are rebuilt unsharded
	WH_XFAIL = 128u,             // a peer's record did not arrive (the host reports the failure)
	WH_XOVER_K4 = 256u           // a rank's record of adjacent pairs overflowed: the pair test of this wave is redone unsharded
};

// Multi-GPU look-ahead (state replicated, one process per GPU): a pass over >= shard_min_rows rows is split by row
// group across the ranks; each rank collects its list entries in a send record, stores the record into every
// peer's receive area over NVLink (peer-mapped memory, cudaIpc) and publishes the pass number in the peer's flag
// word; the merge kernel of each rank waits for the flags and appends all records to its own lists.  No host
// involvement and no collective launch: every rank's scheduler takes the same decisions from the same state.
#define B200_X_CAP 16382u        // entries (u64: slot << 32 | list entry) per record
#define B200_X_WORDS (B200_X_CAP + 2u)   // u64 words per record: [0] = entry count (may exceed the capacity = overflow), [1] = pass number
// The pair test of a wave is split the same way (tile pairs of all cuts dealt round-robin over ranks x blocks): each rank
// tests its share, collects the ADJACENT pairs it finds (wave position << 56 | a << 28 | b) and exchanges them like the
// look-ahead records; the merge kernel files every pair under its cut, so the adjacency build sees the complete list.
#define B200_XK_CAP 65532u       // adjacent pairs per record
#define B200_XK_WORDS (B200_XK_CAP + 4u)  // [0] count, [1] exchange number, [2] bit0: a local survivor list overflowed, [3] survivors it needs
#define B200_X_MAXRANKS 8
#define B200_WV_GROUP 512u         // rows per group of the look-ahead kernel (2 per thread of a 256-thread block): the unit of the rank split
#define ST_WAVE_DEFER 2048u      // (CutCtl::status, wave path) this cut goes back to the pending list untouched

struct WaveCtl {
	// ---- batch description (host-written)
	u32 n_total;                 // halfspaces in the batch
	u32 facet0;                  // dual slot of halfspace 0
	u32 batch_first;
	u32 max_wave;                // <= B200_WAVE_MAXW: clusters the tail kernels are launched with
	u32 cand;                    // pending slots examined when a wave is formed
	u32 in_order;                // 1: a wave is a run of consecutive pending cuts (stops at the first conflict); 0: conflicting cuts are skipped
	u32 refill_below;            // a look-ahead pass runs when fewer slots than this are pending
	// ---- scheduler state
	u32 next_hs;                 // next halfspace without a slot
	u32 done_hs;                 // halfspaces processed
	u32 n_pending;
	u32 pending[B200_WAVE_SLOTS];   // slot ids, ascending halfspace index
	u32 slot_hs[B200_WAVE_SLOTS];   // halfspace held by a slot, B200_NONE = free
	u32 n_la;
	u32 la[B200_WAVE_SLOTS];        // slots the next look-ahead pass classifies
	u32 la_rows;                 // rows that pass covers
	u32 reclassify;              // all pending lists are stale (serial cut, compaction): the next pass rebuilds them
	u32 n_wave;
	u32 wave[B200_WAVE_MAXW];       // slots of the current wave, ascending halfspace index
	u32 n_commit;                // wave positions [0, n_commit) are carried out, the others deferred
	u32 epoch;                   // tag of the footprint marks
	u32 iter;                    // completed wave iterations
	u32 halt, halt_hs, halt_rows, halt_inc, halt_adj, halt_pairs;
	u32 la_new;                  // bit k: entry k of `la` is a halfspace that just received its slot (parameters to be built)
	u32 shard;                   // this look-ahead pass is sharded across the ranks (row groups by rank, exchange over peer memory)
	u64 halt_bits;
	u32 xseq;                    // sequence number of the latest sharded pass (process-wide, identical on every rank)
	u32 noshard_once;            // the next pass runs unsharded (an exchange record overflowed)
	u32 shard_k4;                // the pair test of every wave is split across the ranks (host-set, the same on every rank)
	u32 xseq_k;                  // number of the latest pair-test exchange (process-wide, identical on every rank)
	// ---- statistics (same meaning as EngineStats)
	u64 st_cuts, st_redundant, st_evals, st_rows_scanned, st_minus, st_zero, st_edge, st_copies, st_pair_tests, st_pairs, st_bytes;
	u64 st_waves, st_la_passes, st_deferred, st_sharded, st_sharded_k4;
	u64 t_first, t_last;         // %globaltimer of the first and the latest commit (diagnostics)
};
// The kernels stage WaveCtl in shared memory, let one thread work on the copy and write it back with all threads
// (a single thread walking global memory pays a full round trip per access); what is updated with atomics
// therefore lives outside it (WaveDev::wflag, WaveDev::fin_ctr).
struct WaveCut {                 // what the plan and the commit need of one cut of the wave (gathered by parallel threads)
	u32 status, n_new, inc_new, padj_new, n_minus, n_zero, n_pairs, n_surv, adj_new, live_before, facet, hs, slot;
};

struct WaveProgress {            // mapped pinned host memory, written by the last kernel of an iteration
	volatile u32 iter, done_hs, halt, nrows, n_live, pad[3];
};

struct WaveDev {                 // device pointers of the wave path (kernel argument, by value)
	WaveCtl *wc;
	CutCtl *ctl;                 // [B200_WAVE_SLOTS] per-slot control block: look-ahead accumulators, then the cut's counters and bases
	CutParams *cur;              // [B200_WAVE_SLOTS]
	u32 *list;                   // [B200_WAVE_SLOTS][B200_WAVE_LIST]
	u32 *mark;                   // [cap_rows] footprint marks (epoch << 5 | 31 - pending position)
	int *rc;                     // [n_total] return codes (0 cut, 1 redundant)
	u32 *wflag;                  // [B200_WAVE_SLOTS] wave formation, by pending position: bit0 conflict, bit1 must run alone
	u32 *fin_ctr;                // clusters of the last kernel of an iteration that have finished (the last one commits)
	u64 *trace;                  // [256][8] %globaltimer at the start of each kernel of the last 256 iterations (diagnostics, B200_WAVE_TRACE)
	WaveProgress *progress;
	// scratch of one cut, replicated per wave position
	u32 *vis, *cnt3, *base3, *dead_slots, *he_off, *he_own, *he_inc, *he_k, *he_rank, *he_incpre;
	u8 *he_flag;
	u64 *zmask;
	u32 *padj, *new_padj_off, *new_padj_len, *new_parent, *deg, *adj_fill, *adj_base;
	u32 *pair_a, *pair_b, *surv_a, *surv_b, *facet_epoch, *facet_local, *dead_facets;
	u64 *bits;
	u32 cap_new, cap_pairs, cap_facets;   // per wave position
	u64 cap_bits;
	u32 cap_he;
	// multi-GPU look-ahead exchange (null / 0 on one GPU)
	u32 nranks, rank, shard_min_rows;
	unsigned long long *xsend;                 // [B200_X_WORDS] this rank's record of the current pass
	unsigned long long *xrecv;                 // [nranks][2][B200_X_WORDS] records stored by the peers (double-buffered by pass parity)
	u32 *xflag;                                // [nranks * 32] pass number published by each peer (one 128-byte line each)
	unsigned long long *xpeer_recv[B200_X_MAXRANKS];   // peer-mapped xrecv of every other rank
	u32 *xpeer_flag[B200_X_MAXRANKS];          // peer-mapped xflag of every other rank
	unsigned long long *xksend, *xkrecv;       // [B200_XK_WORDS], [nranks][2][B200_XK_WORDS]: the same for the adjacent pairs of a wave
	u32 *xkflag;
	unsigned long long *xkpeer_recv[B200_X_MAXRANKS];
	u32 *xkpeer_flag[B200_X_MAXRANKS];
};
