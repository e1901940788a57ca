// oracle/cf_oracle.h -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the Centrifuge per-read classification hot path
// (reference: DaehwanKimLab/centrifuge v1.0.4).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
// The product (centrifuge_b200/) never includes, links or executes this code.
//
// Parity status: PINNED.  The restatement is checked byte-for-byte against
//  (a) the reference's only known-answer fixture (example/ + MANUAL.markdown:1586-1603) and
//  (b) the unmodified reference binary compiled into oracle/_ref/ on seeded synthetic inputs
// by tests/test_oracle_*.py.
#ifndef CF_ORACLE_H_
#define CF_ORACLE_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cfo_index cfo_index;

// One classification record, identical in layout to cfb_result_rec of include/cfb200.h.
typedef struct {
	uint64_t taxid;
	uint32_t score;
	uint32_t hitlen;   // (uint64_t)summedHitLen
	uint32_t uid;      // sequence index of the hit, 0xFFFFFFFF if merged by tree traversal
	uint32_t pad;
} cfo_rec;

typedef struct {
	int      khits;            // -k (default 5)
	int      min_hitlen;       // --min-hitlen (default 22)
	int      tree_traverse;    // !--no-traverse
	int      class_rank_slot;  // TaxonomyPathTable::rank_to_pathID(--classification-rank); 0 = strain
	const uint64_t* host_taxids;     size_t n_host;      // as given on the command line
	const uint64_t* excluded_taxids; size_t n_excluded;
} cfo_params;

// Per-batch operation counters used for the roofline's algorithmic-byte definition
// (SURVEY.md section 8d): bytes = 128*S_search + 16*F + 128*S_walk + w*R.
typedef struct {
	uint64_t reads;
	uint64_t partial_searches;   // calls of partialSearch
	uint64_t ftab_probes;        // F
	uint64_t lf_range_steps;     // range steps (top & bot)
	uint64_t lf_range_same_side; // ... of which top and bot share one side
	uint64_t lf_single_steps;    // single-row steps (mapLF1)
	uint64_t sides_search;       // S_search
	uint64_t walk_steps;         // S_walk
	uint64_t rows_resolved;      // R
	uint64_t hits_resolved;      // partial hits sent to resolve()
	uint64_t ext_searches;       // partial searches issued by the extend step
} cfo_stats;

cfo_index* cfo_index_load(const char* basename, char* err, size_t errlen);
void       cfo_index_free(cfo_index*);
uint64_t   cfo_index_len(const cfo_index*);
uint64_t   cfo_index_nseq(const cfo_index*);
int        cfo_index_sample_width(const cfo_index*);   // 2 or 4
int        cfo_index_compressed(const cfo_index*);

// Classify n units.  Unit i is a single read (mate2 length 0 / mate2_off NULL) or a pair.
// bases: 1 byte per base (0..3, 4 = N), forward strand as read from the file.
// flags[i]: bit0 = mate1 passes the host-side filters, bit1 = mate2 passes.
// out_n[i] = number of records for unit i; records are appended to out (capacity cap);
// a unit with zero surviving entries gets one record {taxid 0, score 0, hitlen 0, uid 0xFFFFFFFF}
// flagged by out_n[i] == 0 (no record is written).  Returns total records or <0 on error.
int64_t cfo_classify(const cfo_index*, const cfo_params*,
                     const uint8_t* bases, const uint64_t* off1, const uint32_t* len1,
                     const uint64_t* off2, const uint32_t* len2, const uint8_t* flags,
                     size_t n, uint32_t* out_n, cfo_rec* out, size_t cap, cfo_stats* stats);

// Stage dump for kernel-by-kernel diffs: partial hits of one read after
// searchForwardAndReverse (incl. extension/trim).  Arrays of capacity cap per strand.
// Returns 0; n_hits[0..1] = hits per strand (fw, rc).
int cfo_search_dump(const cfo_index*, const cfo_params*, const uint8_t* bases, uint32_t len,
                    int after_trim, uint32_t n_hits[2], uint64_t* top, uint64_t* bot,
                    uint32_t* bwoff, uint32_t* hlen, size_t cap);

// LF / rank primitives (unit tests of the device primitives).
uint64_t cfo_lf(const cfo_index*, uint64_t row, int c);
int      cfo_bwt_char(const cfo_index*, uint64_t row);
uint64_t cfo_resolve(const cfo_index*, uint64_t row, uint64_t* steps);
void     cfo_ftab_lohi(const cfo_index*, const uint8_t* seq10, uint64_t* top, uint64_t* bot);

// End-to-end file driver (same observable behaviour as `centrifuge-class` for the
// options it understands).  argv-style; returns process exit code.
int cfo_main(int argc, const char** argv);

#ifdef __cplusplus
}
#endif
#endif
