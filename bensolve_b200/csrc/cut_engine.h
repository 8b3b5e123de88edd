// Host-side owner of one polytope's device state and of the per-cut kernel pipeline.
#pragma once
#include <stddef.h>

#include <string>
#include <functional>
#include <vector>

#include "cut_types.h"

struct CutDelta {             // what one cut changed, in host slot numbers (SURVEY 8(b) coherence rule).
	// The arrays point into the engine's pinned staging buffer and stay valid until its next cut.
	int redundant = 0;        // 1 => nothing changed, poly__add_vrtx returns EXIT_FAILURE
	u32 trigger_slot = 0;     // lowest strictly violated slot: the reference's args->idx (bslv_poly.c:121-131)
	u32 first_new_slot = 0;
	u32 n_new = 0;
	const double *coords = nullptr;     // AoS [n_new][d]
	const u8 *ideal = nullptr;          // [n_new]
	const u32 *parent_slot = nullptr;   // [n_new] slot copied from (ZERO copies) or B200_NONE (edge vertices)
	const u32 *dead_slots = nullptr;    // [n_dead_entries], entries equal to B200_NONE are to be skipped
	u32 n_dead_entries = 0;
	const u32 *dead_facets = nullptr;   // [n_dead_facets]
	u32 n_dead_facets = 0;
	bool applied_early = false;         // the on_early callback of cut() has already received this record
};

struct MirrorDump {           // bulk state for rebuilding the host mirror after a device-resident batch
	u32 nrows = 0, slot_cnt = 0;  // (views into the engine's pinned bulk buffer, valid until the next engine call)
	const u32 *row_slot = nullptr, *live_words = nullptr, *ideal_words = nullptr, *root = nullptr, *facet_alive = nullptr;
	const double *coords_soa = nullptr;   // [d][nrows]
};

struct HostStructure {        // snapshot for lazy materialisation of the host poly_lists
	u32 nrows = 0;
	std::vector<u32> row_slot, live_words, inc_off, inc_len, adj_off, adj_len, inc_pool, adj_pool;
};

struct EngineStats {
	u64 cuts = 0, redundant = 0, vertex_evals = 0, rows_scanned = 0;
	u64 minus = 0, zero = 0, zero_plus_projected = 0, edge_vertices = 0, copies = 0;
	u64 pair_tests = 0, new_adjacent_pairs = 0, algorithmic_bytes = 0, kernel_launches = 0, compactions = 0;
	double classify_ms = 0, cut_ms = 0;
	u64 phase_ns[16] = {0};
	u64 sub_ns[16] = {0};
	double host_us[12] = {0};     // [8] device growth (ensure_*), [9] growth events
	u64 redo_loops = 0;
	u64 waves = 0, wave_cuts = 0, la_passes = 0, wave_deferred = 0, wave_serial = 0, wave_halts = 0;   // wave path
	u64 sharded_passes = 0, sharded_cuts = 0, sharded_pair_tests = 0;      // several ranks: look-ahead passes / per-call cuts split across them
};

class CutEngine {
public:
	explicit CutEngine(int dim);
	~CutEngine();
	CutEngine(const CutEngine &) = delete;
	CutEngine &operator=(const CutEngine &) = delete;

	int dim() const { return d_; }
	// Load the start polyhedron (poly__poly_initialise, bslv_poly.c:711-787): n rows = slots 0..n-1.
	void upload_initial(u32 n, const double *coords_aos, const u8 *ideal,
	                    const std::vector<std::vector<u32>> &inc, const std::vector<std::vector<u32>> &adj,
	                    u32 n_facets, const std::vector<u32> &facet_counts);
	// One halfspace; fills `out`.  Throws std::runtime_error on CUDA errors.
	// on_early (optional) is called from inside cut() with the complete delta as soon as the device has written it,
	// while the adjacency build of the same cut still runs; out.applied_early tells the caller it has been called
	void cut(const CutParams &P, CutDelta &out, const std::function<void(const CutDelta &)> *on_early = nullptr);
	// Device-resident batch path: halfspace i of `d_vals` (device memory, [n][d], default callback
	// meaning), no delta transfer; only the 128-byte header comes back.  Returns 1 if redundant.
	int cut_from_device(const double *d_vals, const unsigned char *d_ideal, u64 i, u32 facet, u32 batch_first);
	// The whole batch: halfspaces 0..n-1 of `d_vals` become facets facet0..facet0+n-1.  Small polytopes go through
	// cut_from_device one by one; from a few 10^4 live vertices on the wave path takes over (look-ahead
	// classification, concurrent commuting cuts; cut_types.h "Wave path").  rc_out[i] = 1 if halfspace i was
	// redundant.  Returns the number of cuts.
	long cut_batch_from_device(const double *d_vals, const unsigned char *d_ideal, u64 n, u32 facet0, u32 batch_first, int *rc_out);
	void download_mirror(MirrorDump &out, u32 n_facets);
	void dual_adjacency(const std::vector<u32> &facet_rank, u32 n_live_facets, std::vector<u32> &pair_a, std::vector<u32> &pair_b);
	void reserve(u64 rows, u64 inc_entries, u64 adj_entries);
	// Launch K1 alone `iters` times against halfspace P (no mutation), optionally flushing L2
	// before each launch; returns the mean CUDA-event time of one launch in ms.
	double classify_bench(const CutParams &P, int iters, int flush_l2);
	void *device_alloc(size_t bytes);
	void device_free(void *p);
	void device_upload(void *dst, const void *src, size_t bytes);
	void device_download(void *dst, const void *src, size_t bytes);
	// Overwrite device coordinates of the given live slots from the host mirror (after the caller
	// edited primal.data in place, bslv_algs.c:193-273).
	void reupload_coords(const double *data_aos, size_t n_slots);
	void download_structure(HostStructure &out);
	void compact();                      // drop dead rows, repack pools
	void set_flags(unsigned f) { flags_ = f; }
	const EngineStats &stats() const { return stats_; }
	u32 live_vertices() const { return hdr_.n_live; }
	u32 slots() const { return hdr_.slot_cnt; }
	u32 rows() const { return hdr_.nrows; }

private:
	void drop_shadow();
	bool shard_now() const;          // several ranks: is K1 of the next cut split by row range (large polytopes only)?
	void ensure_rows(u32 need);
	void ensure_inc(u32 need);
	void ensure_adj(u32 need);
	void ensure_padj(u32 need);
	void ensure_pairs(u32 need);
	void ensure_bits(u64 need);
	void ensure_facets(u32 need);
	void launch_part_a(const CutParams &P);
	void launch_classify_dim(int gcls);
	void launch_k1_lists(const CutParams &P, const double *dv, const unsigned char *di, u64 vi, bool sharded);
	void tile_range(bool sharded, u32 &lo, u32 &hi) const;
	void launch_small(const CutParams &P, int mode, bool header_only);
	void launch_k4_and_tail2(bool header_only);
	bool use_small_path() const;
	bool wide_cluster() const;
	bool ensure_shadow();
	void run_cut(const CutParams &P, bool header_only);
	void account(const CutParams &P, u32 n_live_before, u32 nrows_before);
	void launch_part_b(bool rerun);
	void launch_part_c(bool header_only);
	void fetch_delta();
	void bump_seq();
	void ensure_stage(u64 need);
	void maybe_compact();
	// wave path
	void wave_ensure_scratch(u32 n_facets, u32 pairs_per_pos, u64 bits_per_pos);
	void wave_free();
	bool wave_adopt();                      // take over the parked wave scratch of a killed polytope, if any
	void wave_enqueue(int from_stage, const double *d_vals, const unsigned char *d_ideal);
	void wave_sync_ctl(WaveCtl &wc);        // stream idle; host copies of the wave and the main control block
	void wave_upload_ctl(const WaveCtl &wc);

	int d_;
	unsigned flags_ = 0;
	bool header_only_ = false;
	u32 seq_ = 0;                  // sequence number of the record the host waits for
	bool tiny_caps_ = false;
	bool small_dirty_ = true;      // tile counters / K1 accumulators must be cleared before the small-cut path runs
	bool prefer_big_ = false;      // the last cut did not fit the single-CTA tail
	u32 emu_extra_status_ = 0;
	u32 expect_vis_ = 0;
	u32 expect_m_ = 0;             // new vertices of the previous cut (predicts whether the tail can run K4 itself)
	DevState S_{};
	CutCtl hdr_{};            // host copy of the control block as of the last sync
	CutCtl *pinned_hdr_ = nullptr;
	unsigned char *pinned_stage_ = nullptr;
	const std::function<void(const CutDelta &)> *on_early_ = nullptr;
	bool early_done_ = false;
	u32 early_n_new_ = 0;
	u32 *gc_totals_ = nullptr;               // device scratch of compact()
	unsigned char *pinned_bulk_ = nullptr;   // staging for bulk downloads (mirror rebuild), grown geometrically
	size_t pinned_bulk_cap_ = 0;
	void *stream_ = nullptr;
	void *ev_[4] = {nullptr, nullptr, nullptr, nullptr};
	int num_sms_ = 148;
	int nranks_ = 1, rank_ = 0;          // communicator at construction time
	const double *dev_vals_ = nullptr;      // set while a device-resident batch is running
	const unsigned char *dev_ideal_ = nullptr;
	u64 dev_index_ = 0;
	void *flush_buf_ = nullptr;
	void *row_scratch_ = nullptr;       // one block behind the per-cut scratch arrays indexed by row (S_.cls .. S_.tile_base)
	void *shadow_[11] = {nullptr};      // second set of persistent arrays: target of the next compaction
	bool shadow_valid_ = false;
	u32 shadow_rows_ = 0, shadow_inc_ = 0, shadow_adj_ = 0;
	WaveDev WD_{};                 // wave path scratch (allocated on first use)
	u32 wave_rows_ = 0;            // rows the mark array covers
	WaveProgress *wave_progress_ = nullptr;   // mapped pinned host memory (device alias in WD_.progress)
	u32 wave_epoch_ = 1;
	u64 wave_rc_cap_ = 0;
	int wave_max_clusters_ = 0;    // co-resident clusters of the per-cut wave kernels (0 = not queried yet)
	EngineStats stats_;
};

// multi-GPU communicator (cut_engine.cu); the callback form exists in the host test double only
typedef void (*b200_allgather_fn)(const void *send, void *recv, size_t bytes_per_rank);
int b200_comm_make_id(char out[128]);
int b200_comm_start(int rank, int nranks, const char id_bytes[128]);
void b200_comm_stop();
int b200_comm_set_callback(b200_allgather_fn fn);
int b200_comm_rank();
int b200_comm_size();

// library-wide helpers (cut_engine.cu)
void b200_set_error(const std::string &msg);
const char *b200_get_error();
int b200_select_device(int dev);
int b200_num_devices();
