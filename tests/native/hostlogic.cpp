// tests/native/hostlogic.cpp -- TEST ONLY.  Compiles the product's scalar classification logic
// (centrifuge_b200/csrc/cf_logic.h, the code the CUDA kernels run per thread) for the host and
// exposes it with the oracle's cfo_classify signature so tests can diff the two record streams.
// This is not a CPU fallback: nothing in the product links or loads this file.
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <set>
#include "../../centrifuge_b200/csrc/cf_index.h"
#include "../../centrifuge_b200/csrc/cf_logic.h"

using namespace cfb;

struct HL { HostIndex h; IndexView v; std::vector<uint8_t> excl; std::vector<uint64_t> host; };

static void expand(const HostIndex& h, const uint64_t* ids, uint64_t n, std::set<uint64_t>& out) {
	if(!n) return;
	for(size_t i = 0; i < h.nodes.size(); i++) {
		uint64_t t = h.nodes[i].taxid;
		for(;;) {
			bool f = false; for(uint64_t k = 0; k < n; k++) if(ids[k] == t) f = true;
			if(f) { out.insert(h.nodes[i].taxid); break; }
			const TaxNode* nd = h.find_node(t);
			if(!nd || nd->parent == t) break;
			t = nd->parent;
		}
	}
}

extern "C" void* hl_load(const char* base, char* err, size_t errlen) {
	HL* x = new HL();
	std::string e = load_cf_index(base, x->h);
	if(!e.empty()) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; delete x; return NULL; }
	if(x->h.line_rate != 7) { strncpy(err, "the scalar logic (like the kernels) assumes 128-byte sides: lineRate 7 only", errlen - 1); err[errlen - 1] = 0; delete x; return NULL; }
	const HostIndex& h = x->h; IndexView& v = x->v; memset(&v, 0, sizeof v);
	v.sides = (const uint64_t*)h.sides.data(); v.ftab = h.ftab.data(); v.eftab = h.eftab.data();
	v.sample16 = h.wide_sample ? NULL : h.sample16.data(); v.sample32 = h.wide_sample ? h.sample32.data() : NULL;
	v.brow = h.brow.data(); v.bseq = h.bseq.data(); v.bbits = h.bbits.data();
	v.seq_taxid = h.seq_taxid.data(); v.seq_path = h.seq_path.data(); v.paths = h.paths.data();
	v.len = h.len; v.zoff = h.zoff; v.zside = h.zoff / 384; v.zoffc = (uint32_t)(h.zoff % 384);
	for(int i = 0; i < 4; i++) v.fchr[i] = h.fchr[i];
	v.last_boundary = h.last_boundary; v.num_sides = h.num_sides; v.n_boundaries = (uint32_t)h.brow.size();
	v.n_seqs = (uint32_t)h.seq_taxid.size(); v.off_rate = h.off_rate; v.ftab_chars = h.ftab_chars; v.bshift = h.bshift;
	return x;
}
extern "C" void hl_free(void* p) { delete (HL*)p; }
extern "C" int hl_line_rate(void* p) { return ((HL*)p)->h.line_rate; }

struct OParams { int khits, min_hitlen, tree_traverse, class_rank_slot; const uint64_t* host; size_t n_host; const uint64_t* excl; size_t n_excl; };

extern "C" long long hl_classify(void* hp, const OParams* op, const uint8_t* bases, const uint64_t* off1, const uint32_t* len1,
                                 const uint64_t* off2, const uint32_t* len2, const uint8_t* flags, size_t n,
                                 uint32_t* out_n, OutRec* out, size_t cap, unsigned long long* counters /*8 or NULL*/) {
	HL* x = (HL*)hp; const HostIndex& h = x->h;
	Params p; p.khits = op->khits; p.min_hitlen = op->min_hitlen < 15 ? 15 : op->min_hitlen;
	p.ihits = (uint32_t)std::max(op->khits, 5) * (h.compressed ? 4u : 40u);
	p.increment = (2 * p.min_hitlen <= 33) ? 10 : (2 * p.min_hitlen - 33);
	p.tree_traverse = op->tree_traverse; p.class_rank_slot = (uint32_t)op->class_rank_slot & 0xff;
	std::set<uint64_t> hs, es; expand(h, op->host, op->n_host, hs); expand(h, op->excl, op->n_excl, es);
	IndexView v = x->v;
	x->excl.assign(h.seq_taxid.size(), 0);
	if(!es.empty()) { for(size_t i = 0; i < x->excl.size(); i++) x->excl[i] = es.count(h.seq_taxid[i]) ? 1 : 0; v.seq_excluded = x->excl.data(); }
	x->host.assign(hs.begin(), hs.end());
	if(!x->host.empty()) { v.host_taxids = x->host.data(); v.n_host = (uint32_t)x->host.size(); }
	Counters ctr; memset(&ctr, 0, sizeof ctr);
	size_t total = 0;
	for(size_t i = 0; i < n; i++) {
		uint8_t fl = flags ? flags[i] : 1;
		const bool pair = off2 && len2 && (fl & 4);
		if(!pair) fl &= 1;
		UnitHits u; u.n_mates = 0; const uint8_t* fw[2];
		std::vector<HitRec> store[2][2];
		for(int m = 0; m < (pair ? 2 : 1); m++) {
			if(!((fl >> m) & 1)) continue;
			const uint32_t len = m == 0 ? len1[i] : len2[i];
			if(len == 0) continue;
			const int r = u.n_mates++;
			fw[r] = bases + (m == 0 ? off1[i] : off2[i]); u.rdlen[r] = len;
			const uint32_t capn = len + 2;
			for(int s = 0; s < 2; s++) {
				store[r][s].resize(capn);
				u.n[r][s] = search_strand_scalar(v, p, fw[r], len, s, store[r][s].data(), capn, &ctr);
				u.L[r][s] = store[r][s].data();
			}
		}
		uint32_t no = 0;
		std::vector<OutRec> recs;
		if(u.n_mates > 0) {
			ctr.units++;
			for(int r = 0; r < u.n_mates; r++) post_search(v, p, fw[r], u.rdlen[r], u.L[r][0], u.n[r][0], u.L[r][1], u.n[r][1], &ctr);
			SortAndCount sc(p, u); for_each_visit(p, u, sc);
			std::vector<uint64_t> rows(sc.rows + 1); std::vector<uint32_t> ids(sc.rows + 1);
			EmitRows er(p, u, rows.data()); for_each_visit(p, u, er);
			if(er.k != sc.rows) return -2;
			for(uint64_t k = 0; k < sc.rows; k++) { ids[k] = resolve_scalar(v, rows[k] & kRowMask, &ctr); ctr.rows_resolved++; }
			std::vector<Entry> ent(sc.rows + 1); std::vector<TaxCnt> tc(sc.rows + 1); recs.resize(sc.rows + 1);
			{ CountRows cr(p, u); for_each_visit(p, u, cr); if(cr.rows != sc.rows) return -3; }     // the re-run path counts sorted lists
			const uint32_t nmap = score_plan(v, p, rows.data(), ids.data(), sc.rows, ent.data());
			no = reduce_and_emit(v, p, u.n_mates == 2, ent.data(), nmap, tc.data(), recs.data());
		}
		out_n[i] = no;
		if(total + no > cap) return -1;
		for(uint32_t k = 0; k < no; k++) out[total++] = recs[k];
	}
	if(counters) { counters[0] = ctr.units; counters[1] = ctr.partial_searches; counters[2] = ctr.ftab_probes; counters[3] = ctr.sides_search;
		counters[4] = ctr.walk_steps; counters[5] = ctr.rows_resolved; counters[6] = ctr.lf_steps; counters[7] = ctr.ext_searches; }
	return (long long)total;
}

// std::sort twin check: sorts (len,size) pairs with the product's restated introsort; the test
// compares against libstdc++'s std::sort on the same data.
extern "C" void hl_sort_hits(HitRec* h, size_t n) { std_sort(h, h + n, HitLess()); }
extern "C" void hl_std_sort_hits(HitRec* h, size_t n) { std::sort(h, h + n, HitLess()); }
