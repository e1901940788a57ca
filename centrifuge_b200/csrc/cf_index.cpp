// cf_index.cpp -- parse `.1-.4.cf` into flat host arrays (see cf_index.h for format citations).
#include "cf_index.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <set>

namespace cfb {

static const char* const kRank[RANK_MAX] = {
	"no rank", "strain", "species", "genus", "family", "order", "class", "phylum", "kingdom", NULL,
	"forma", "infraclass", "infraorder", "parvorder", "subclass", "subfamily", "subgenus", "subkingdom",
	"suborder", "subphylum", "subspecies", "subtribe", "superclass", "superfamily", "superkingdom",
	"superorder", "superphylum", "tribe", "varietas", "life"};

// RANK_DOMAIN has no spelling in the reference's tables: it prints as "no rank" and is not parsed.
const char* rank_name(int r) { return (r > 0 && r < RANK_MAX && kRank[r]) ? kRank[r] : "no rank"; }
int rank_from_name(const char* s) {
	for(int r = 1; r < RANK_MAX; r++) if(kRank[r] && strcmp(s, kRank[r]) == 0) return r;
	return RANK_UNKNOWN;
}
int rank_to_slot(int rank) {
	switch(rank) {
		case RANK_STRAIN: case RANK_SUB_SPECIES: return 0;
		case RANK_SPECIES: return 1;   case RANK_GENUS: return 2;  case RANK_FAMILY: return 3;
		case RANK_ORDER: return 4;     case RANK_CLASS: return 5;  case RANK_PHYLUM: return 6;
		case RANK_KINGDOM: return 7;   case RANK_SUPER_KINGDOM: return 8; case RANK_DOMAIN: return 9;
		default: return 255;
	}
}
// initial_tax_rank_num taxonomy.h:161-200: coarse level of each rank
static int rank_level(int r) {
	switch(r) {
		case RANK_SUB_SPECIES: case RANK_STRAIN: return 0;
		case RANK_SPECIES: return 1;
		case RANK_SUB_GENUS: case RANK_GENUS: return 2;
		case RANK_SUB_FAMILY: case RANK_FAMILY: case RANK_SUPER_FAMILY: return 3;
		case RANK_SUB_ORDER: case RANK_INFRA_ORDER: case RANK_PARV_ORDER: case RANK_ORDER: case RANK_SUPER_ORDER: return 4;
		case RANK_INFRA_CLASS: case RANK_SUB_CLASS: case RANK_CLASS: case RANK_SUPER_CLASS: return 5;
		case RANK_SUB_PHYLUM: case RANK_PHYLUM: case RANK_SUPER_PHYLUM: return 6;
		case RANK_SUB_KINGDOM: case RANK_KINGDOM: case RANK_SUPER_KINGDOM: return 7;
		case RANK_DOMAIN: case RANK_FORMA: case RANK_SUB_TRIBE: case RANK_TRIBE: case RANK_VARIETAS: case RANK_UNKNOWN: return 8;
		default: return 0;   // RANK_LIFE is never assigned by the reference (static zero)
	}
}

const TaxNode* HostIndex::find_node(uint64_t taxid) const {
	size_t lo = 0, hi = nodes.size();
	while(lo < hi) { size_t mid = (lo + hi) >> 1; if(nodes[mid].taxid < taxid) lo = mid + 1; else hi = mid; }
	return (lo < nodes.size() && nodes[lo].taxid == taxid) ? &nodes[lo] : NULL;
}

namespace {
struct In {
	FILE* f; bool ok;
	explicit In(const std::string& p) : f(fopen(p.c_str(), "rb")), ok(f != NULL) {}
	~In() { if(f) fclose(f); }
	template <typename T> T rd() { T v = 0; if(!ok || fread(&v, sizeof(T), 1, f) != 1) ok = false; return v; }
	void bulk(void* dst, size_t n) {
		char* p = (char*)dst;
		while(ok && n) { size_t k = fread(p, 1, std::min<size_t>(n, (size_t)1 << 30), f); if(k == 0) { ok = false; break; } p += k; n -= k; }
	}
	int ch() { return ok ? fgetc(f) : EOF; }
};
}  // namespace

std::string load_cf_index(const std::string& base, HostIndex& ix, bool defer_bulk) {
	ix.bulk_deferred = defer_bulk;
	// ---------------- .1.cf
	{
		In in(base + ".1.cf");
		if(!in.ok) return "could not open index file " + base + ".1.cf";
		uint32_t one = in.rd<uint32_t>();
		if(one != 1) return "index " + base + ".1.cf has foreign endianness or is not a .cf index";
		ix.len = in.rd<uint64_t>();
		ix.line_rate = in.rd<int32_t>(); (void)in.rd<int32_t>();
		ix.off_rate = in.rd<int32_t>(); ix.ftab_chars = in.rd<int32_t>();
		int32_t flags = in.rd<int32_t>(); (void)flags;
		if(!in.ok) return "truncated header in " + base + ".1.cf";
		if(ix.line_rate < 6 || ix.line_rate > 12 || ix.off_rate < 0 || ix.off_rate > 30 || ix.ftab_chars < 1 || ix.ftab_chars > 15)
			return "implausible header in " + base + ".1.cf";
		ix.bwt_len = ix.len + 1;
		ix.side_sz = (uint64_t)1 << ix.line_rate;
		ix.side_bwt_sz = ix.side_sz - 4 * sizeof(uint64_t);
		ix.side_bwt_len = ix.side_bwt_sz * 4;
		ix.num_sides = (ix.len / 4 + 1 + ix.side_bwt_sz - 1) / ix.side_bwt_sz;
		ix.ftab_len = ((uint64_t)1 << (2 * ix.ftab_chars)) + 1;
		ix.eftab_len = 2 * (uint64_t)ix.ftab_chars;
		ix.offs_len = (ix.bwt_len + ((uint64_t)1 << ix.off_rate) - 1) >> ix.off_rate;
		ix.n_pat = in.rd<uint64_t>();
		if(fseeko(in.f, (off_t)(ix.n_pat * 8), SEEK_CUR) != 0) in.ok = false;      // plen[] unused on this path
		uint64_t n_frag = in.rd<uint64_t>();
		if(fseeko(in.f, (off_t)(n_frag * 24), SEEK_CUR) != 0) in.ok = false;       // rstarts[] unused
		ix.wide_sample = ix.n_pat > 65535;
		if(defer_bulk) { ix.sides_file_off = (uint64_t)ftello(in.f); if(fseeko(in.f, (off_t)(ix.num_sides * ix.side_sz), SEEK_CUR) != 0) in.ok = false; }
		else { ix.sides.resize(ix.num_sides * ix.side_sz); in.bulk(ix.sides.data(), ix.sides.size()); }
		ix.zoff = in.rd<uint64_t>();
		for(int i = 0; i < 5; i++) ix.fchr[i] = in.rd<uint64_t>();
		ix.ftab.resize(ix.ftab_len);   in.bulk(ix.ftab.data(), ix.ftab_len * 8);
		ix.eftab.resize(ix.eftab_len); in.bulk(ix.eftab.data(), ix.eftab_len * 8);
		if(!in.ok) return "truncated " + base + ".1.cf";
	}
	// ---------------- .2.cf
	{
		In in(base + ".2.cf");
		if(!in.ok) return "could not open index file " + base + ".2.cf";
		(void)in.rd<uint32_t>();
		if(defer_bulk) {
			ix.sample_file_off = (uint64_t)ftello(in.f);
			if(fseeko(in.f, 0, SEEK_END) != 0 || (uint64_t)ftello(in.f) < ix.sample_file_off + ix.offs_len * (ix.wide_sample ? 4 : 2)) in.ok = false;
		}
		else if(ix.wide_sample) { ix.sample32.resize(ix.offs_len); in.bulk(ix.sample32.data(), ix.offs_len * 4); }
		else                    { ix.sample16.resize(ix.offs_len); in.bulk(ix.sample16.data(), ix.offs_len * 2); }
		if(!in.ok) return "truncated " + base + ".2.cf";
	}
	// ---------------- .3.cf
	{
		In in(base + ".3.cf");
		if(!in.ok) return "could not open index file " + base + ".3.cf";
		(void)in.rd<uint32_t>();
		uint64_t nref = in.rd<uint64_t>();
		std::set<uint64_t> leaf_tids;
		size_t cids = 0;
		for(uint64_t i = 0; i < nref && in.ok; i++) {
			// uid: the reference extracts chars with `istream >> char`, which drops whitespace
			std::string uid;
			for(;;) { int c = in.ch(); if(c == EOF || c == 0) break; if(!isspace(c)) uid.push_back((char)c); }
			uint64_t tid = in.rd<uint64_t>();
			if(uid.compare(0, 3, "cid") == 0) cids++;
			ix.seq_name.push_back(uid); ix.seq_taxid.push_back(tid); leaf_tids.insert(tid);
		}
		ix.compressed = cids >= 10;
		uint64_t ntid = in.rd<uint64_t>();
		std::map<uint64_t, TaxNode> tree;
		while(ntid > 0 && in.ok && tree.size() < ntid) {
			TaxNode n; n.taxid = in.rd<uint64_t>(); n.parent = in.rd<uint64_t>(); n.rank = (uint8_t)in.rd<uint16_t>();
			if(!in.ok) break;
			n.leaf = leaf_tids.count(n.taxid) ? 1 : 0;
			tree[n.taxid] = n;
		}
		for(std::map<uint64_t, TaxNode>::const_iterator it = tree.begin(); it != tree.end(); ++it) ix.nodes.push_back(it->second);
		uint64_t nname = in.rd<uint64_t>();
		while(nname > 0 && in.ok && ix.names.size() < nname) {
			uint64_t tid = in.rd<uint64_t>();
			if(!in.ok) break;
			std::string nm; int c = in.ch();
			while(c != EOF && isspace(c)) c = in.ch();
			while(c != EOF && !isspace(c)) { nm.push_back((char)c); c = in.ch(); }   // delimiter consumed
			std::replace(nm.begin(), nm.end(), '@', ' ');
			ix.names[tid] = nm;
		}
		uint64_t nsize = in.rd<uint64_t>();
		while(nsize > 0 && in.ok && ix.sizes.size() < nsize) {
			uint64_t tid = in.rd<uint64_t>(), sz = in.rd<uint64_t>();
			if(!in.ok) break;
			ix.sizes[tid] = sz;
		}
		// genome size of an internal node = mean over its leaf-level descendants (bt2_idx.h:709-744)
		std::map<uint64_t, uint64_t> cnt, sum;
		for(std::map<uint64_t, uint64_t>::const_iterator it = ix.sizes.begin(); it != ix.sizes.end(); ++it) {
			const TaxNode* n = ix.find_node(it->first);
			if(!n || n->parent == n->taxid) continue;
			if(!((n->rank == RANK_UNKNOWN && n->leaf) || rank_level(n->rank) < rank_level(RANK_SPECIES))) continue;
			uint64_t t = n->parent;
			for(;;) {
				const TaxNode* a = ix.find_node(t);
				if(!a) break;
				if(a->rank == RANK_SPECIES || a->rank == RANK_GENUS || a->rank == RANK_FAMILY ||
				   a->rank == RANK_ORDER || a->rank == RANK_CLASS || a->rank == RANK_PHYLUM) { sum[t] += it->second; cnt[t]++; }
				if(a->parent == t) break;
				t = a->parent;
			}
		}
		for(std::map<uint64_t, uint64_t>::const_iterator it = cnt.begin(); it != cnt.end(); ++it) ix.sizes[it->first] = sum[it->first] / it->second;
		// rank paths (TaxonomyPathTable::buildPaths taxonomy.h:96-149): one per distinct leaf taxid in the tree
		std::map<uint64_t, int32_t> pid;
		ix.seq_path.assign(ix.seq_taxid.size(), -1);
		for(size_t i = 0; i < ix.seq_taxid.size(); i++) {
			uint64_t tid = ix.seq_taxid[i];
			std::map<uint64_t, int32_t>::const_iterator f = pid.find(tid);
			if(f != pid.end()) { ix.seq_path[i] = f->second; continue; }
			if(!ix.find_node(tid)) continue;
			int32_t id = (int32_t)(ix.paths.size() / kPathSlots);
			pid[tid] = id; ix.seq_path[i] = id;
			ix.paths.resize(ix.paths.size() + kPathSlots, 0);
			uint64_t* path = &ix.paths[(size_t)id * kPathSlots];
			bool first = true;
			for(uint64_t t = tid;;) {
				const TaxNode* n = ix.find_node(t);
				if(!n) break;
				int slot = (first && n->rank == RANK_UNKNOWN) ? 0 : rank_to_slot(n->rank);
				if(slot < kPathSlots && path[slot] == 0) path[slot] = t;
				first = false;
				if(n->parent == t) break;
				t = n->parent;
			}
		}
	}
	// ---------------- .4.cf (optional; absent => no boundary rows)
	{
		In in(base + ".4.cf");
		if(in.ok) {
			(void)in.rd<uint32_t>();
			uint64_t n = in.rd<uint64_t>();
			std::map<uint64_t, uint32_t> m;                      // later duplicates overwrite, like the reference's map
			for(uint64_t i = 0; i < n && in.ok; i++) { uint64_t row = in.rd<uint64_t>(); uint32_t s = in.rd<uint32_t>(); if(in.ok) m[row] = s; }
			for(std::map<uint64_t, uint32_t>::const_iterator it = m.begin(); it != m.end(); ++it) { ix.brow.push_back(it->first); ix.bseq.push_back(it->second); }
			if(!ix.brow.empty()) {
				ix.last_boundary = ix.brow.back();
				// prefilter bitmap: ~64 blocks per boundary row
				uint64_t blocks = 64; while(blocks < 64 * (uint64_t)ix.brow.size()) blocks <<= 1;
				ix.bshift = 0; while((((ix.last_boundary) >> ix.bshift) + 1) > blocks) ix.bshift++;
				ix.bbits.assign(((ix.last_boundary >> ix.bshift) + 32) / 32, 0);
				for(size_t i = 0; i < ix.brow.size(); i++) { uint64_t b = ix.brow[i] >> ix.bshift; ix.bbits[b >> 5] |= 1u << (b & 31); }
			}
		}
	}
	return "";
}

}  // namespace cfb
