// oracle/cf_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see cf_oracle.h).
//
// Plain single-threaded C++ restatement of Centrifuge's classification hot path.  Every
// function cites the reference file:line it follows (paths relative to the reference
// tree, DaehwanKimLab/centrifuge v1.0.4).  Written for clarity, not speed; it is the
// executable spec the CUDA kernels are diffed against and the "port" CPU baseline.
#include "cf_oracle.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

typedef uint64_t u64;
typedef uint32_t u32;
static const u64 OFF = 0xffffffffffffffffull;

// ---------------------------------------------------------------------------
// taxonomy.h:16-48 rank enum, :161-200 tax_rank_num, :207-240 rank strings
// ---------------------------------------------------------------------------
enum {
	RANK_UNKNOWN = 0, RANK_STRAIN, RANK_SPECIES, RANK_GENUS, RANK_FAMILY, RANK_ORDER, RANK_CLASS,
	RANK_PHYLUM, RANK_KINGDOM, RANK_DOMAIN, RANK_FORMA, RANK_INFRA_CLASS, RANK_INFRA_ORDER,
	RANK_PARV_ORDER, RANK_SUB_CLASS, RANK_SUB_FAMILY, RANK_SUB_GENUS, RANK_SUB_KINGDOM,
	RANK_SUB_ORDER, RANK_SUB_PHYLUM, RANK_SUB_SPECIES, RANK_SUB_TRIBE, RANK_SUPER_CLASS,
	RANK_SUPER_FAMILY, RANK_SUPER_KINGDOM, RANK_SUPER_ORDER, RANK_SUPER_PHYLUM, RANK_TRIBE,
	RANK_VARIETAS, RANK_LIFE, RANK_MAX
};
static const char* kRankNames[RANK_MAX] = {
	"no rank", "strain", "species", "genus", "family", "order", "class", "phylum", "kingdom",
	"no rank" /*domain has no string in get_tax_rank_string*/, "forma", "infraclass", "infraorder",
	"parvorder", "subclass", "subfamily", "subgenus", "subkingdom", "suborder", "subphylum",
	"subspecies", "subtribe", "superclass", "superfamily", "superkingdom", "superorder",
	"superphylum", "tribe", "varietas", "life"
};
static const char* rank_string(uint8_t r) { return r < RANK_MAX ? kRankNames[r] : "no rank"; }
static uint8_t rank_id(const char* s) {          // taxonomy.h:242-303
	for(int r = 1; r < RANK_MAX; r++) {
		if(r == RANK_DOMAIN) continue;             // "domain" is not parsed by get_tax_rank_id
		if(strcmp(s, kRankNames[r]) == 0) return (uint8_t)r;
	}
	return RANK_UNKNOWN;
}
static uint8_t rank_to_pathID(uint8_t rank) {    // taxonomy.h:68-94
	switch(rank) {
		case RANK_STRAIN: case RANK_SUB_SPECIES: return 0;
		case RANK_SPECIES: return 1; case RANK_GENUS: return 2; case RANK_FAMILY: return 3;
		case RANK_ORDER: return 4; case RANK_CLASS: return 5; case RANK_PHYLUM: return 6;
		case RANK_KINGDOM: return 7; case RANK_SUPER_KINGDOM: return 8; case RANK_DOMAIN: return 9;
		default: return 255;
	}
}
static void fill_tax_rank_num(uint8_t* t) {      // taxonomy.h:161-200
	memset(t, 0, RANK_MAX);
	uint8_t rank = 0;
	t[RANK_SUB_SPECIES] = rank; t[RANK_STRAIN] = rank++;
	t[RANK_SPECIES] = rank++;
	t[RANK_SUB_GENUS] = rank; t[RANK_GENUS] = rank++;
	t[RANK_SUB_FAMILY] = rank; t[RANK_FAMILY] = rank; t[RANK_SUPER_FAMILY] = rank++;
	t[RANK_SUB_ORDER] = rank; t[RANK_INFRA_ORDER] = rank; t[RANK_PARV_ORDER] = rank; t[RANK_ORDER] = rank; t[RANK_SUPER_ORDER] = rank++;
	t[RANK_INFRA_CLASS] = rank; t[RANK_SUB_CLASS] = rank; t[RANK_CLASS] = rank; t[RANK_SUPER_CLASS] = rank++;
	t[RANK_SUB_PHYLUM] = rank; t[RANK_PHYLUM] = rank; t[RANK_SUPER_PHYLUM] = rank++;
	t[RANK_SUB_KINGDOM] = rank; t[RANK_KINGDOM] = rank; t[RANK_SUPER_KINGDOM] = rank++;
	t[RANK_DOMAIN] = rank; t[RANK_FORMA] = rank; t[RANK_SUB_TRIBE] = rank; t[RANK_TRIBE] = rank;
	t[RANK_VARIETAS] = rank; t[RANK_UNKNOWN] = rank;
}

struct TaxNode { u64 parent; uint8_t rank; uint8_t leaf; };

// ---------------------------------------------------------------------------
// Index: bt2_io.h:42-685 (.1.cf/.2.cf), bt2_idx.h:566-854 (.3.cf/.4.cf), :133-167 geometry
// ---------------------------------------------------------------------------
struct cfo_index {
	u64 len = 0; int lineRate = 0, offRate = 0, ftabChars = 0;
	u64 nPat = 0;
	std::vector<u64> plen;
	std::vector<uint8_t> ebwt;
	u64 zOff = 0; u64 fchr[5] = {0, 0, 0, 0, 0};
	std::vector<u64> ftab, eftab;
	bool offw = false;
	std::vector<uint16_t> offs16; std::vector<u32> offs32;
	u64 sideSz = 0, sideBwtSz = 0, sideBwtLen = 0, numSides = 0, ftabLen = 0, eftabLen = 0,
	    offsLen = 0, offMask = 0, bwtLen = 0;
	std::vector<std::pair<std::string, u64> > uid_to_tid;
	std::map<u64, TaxNode> tree;
	std::map<u64, std::string> name;
	std::map<u64, u64> size;
	bool compressed = false;
	std::map<u64, u32> tid_to_pid; std::vector<std::vector<u64> > paths;
	std::map<u64, u32> boundary; u64 lastBoundary = 0;
};

namespace {

struct Reader {
	FILE* f; bool ok;
	explicit Reader(const std::string& p) : f(fopen(p.c_str(), "rb")), ok(f != NULL) {}
	~Reader() { if(f) fclose(f); }
	template <typename T> T get() { T v = 0; if(fread(&v, sizeof(T), 1, f) != 1) ok = false; return v; }
	void bulk(void* dst, size_t n) { if(n && fread(dst, 1, n, f) != n) ok = false; }
	int ch() { int c = fgetc(f); return c; }
};

void build_paths(cfo_index& ix) {                // taxonomy.h:96-149
	ix.tid_to_pid.clear(); ix.paths.clear();
	for(size_t i = 0; i < ix.uid_to_tid.size(); i++) {
		u64 tid = ix.uid_to_tid[i].second;
		if(ix.tid_to_pid.count(tid)) continue;
		if(!ix.tree.count(tid)) continue;
		ix.tid_to_pid[tid] = (u32)ix.paths.size();
		ix.paths.push_back(std::vector<u64>(10, 0));
		std::vector<u64>& path = ix.paths.back();
		bool first = true;
		while(true) {
			std::map<u64, TaxNode>::const_iterator itr = ix.tree.find(tid);
			if(itr == ix.tree.end()) break;
			const TaxNode& node = itr->second;
			u32 rank = 0xffffffffu;
			if(first && node.rank == RANK_UNKNOWN) rank = 0;
			else { uint8_t p = rank_to_pathID(node.rank); if(p != 255) rank = p; }
			if(rank < path.size() && path[rank] == 0) path[rank] = tid;
			first = false;
			if(node.parent == tid) break;
			tid = node.parent;
		}
	}
}

bool load_index(cfo_index& ix, const std::string& base, std::string& err) {
	{
		Reader r(base + ".1.cf");
		if(!r.ok) { err = "cannot open " + base + ".1.cf"; return false; }
		if(r.get<u32>() != 1) { err = "bad endianness sentinel in .1.cf"; return false; }
		ix.len = r.get<u64>(); ix.lineRate = r.get<int32_t>(); (void)r.get<int32_t>();
		ix.offRate = r.get<int32_t>(); ix.ftabChars = r.get<int32_t>(); (void)r.get<int32_t>();
		// EbwtParams::init bt2_idx.h:133-167
		ix.bwtLen = ix.len + 1;
		u64 bwtSz = ix.len / 4 + 1;
		ix.sideSz = (u64)1 << ix.lineRate; ix.sideBwtSz = ix.sideSz - 32; ix.sideBwtLen = ix.sideBwtSz * 4;
		ix.numSides = (bwtSz + ix.sideBwtSz - 1) / ix.sideBwtSz;
		ix.ftabLen = ((u64)1 << (ix.ftabChars * 2)) + 1; ix.eftabLen = (u64)ix.ftabChars * 2;
		ix.offsLen = (ix.bwtLen + ((u64)1 << ix.offRate) - 1) >> ix.offRate;
		ix.offMask = OFF << ix.offRate;
		ix.nPat = r.get<u64>();
		ix.plen.resize(ix.nPat); r.bulk(ix.plen.data(), ix.nPat * 8);
		ix.offw = ix.nPat > 65535;                 // bt2_io.h:280
		u64 nFrag = r.get<u64>();
		if(fseeko(r.f, (off_t)(nFrag * 24), SEEK_CUR) != 0) r.ok = false;
		ix.ebwt.resize(ix.numSides * ix.sideSz); r.bulk(ix.ebwt.data(), ix.ebwt.size());
		ix.zOff = r.get<u64>();
		for(int i = 0; i < 5; i++) ix.fchr[i] = r.get<u64>();
		ix.ftab.resize(ix.ftabLen); r.bulk(ix.ftab.data(), ix.ftabLen * 8);
		ix.eftab.resize(ix.eftabLen); r.bulk(ix.eftab.data(), ix.eftabLen * 8);
		if(!r.ok) { err = "short read in .1.cf"; return false; }
	}
	{
		Reader r(base + ".2.cf");
		if(!r.ok) { err = "cannot open " + base + ".2.cf"; return false; }
		(void)r.get<u32>();
		if(ix.offw) { ix.offs32.resize(ix.offsLen); r.bulk(ix.offs32.data(), ix.offsLen * 4); }
		else        { ix.offs16.resize(ix.offsLen); r.bulk(ix.offs16.data(), ix.offsLen * 2); }
		if(!r.ok) { err = "short read in .2.cf"; return false; }
	}
	{   // bt2_idx.h:623-707 (istream >> char skips whitespace; >> string stops at whitespace)
		Reader r(base + ".3.cf");
		if(!r.ok) { err = "cannot open " + base + ".3.cf"; return false; }
		(void)r.get<u32>();
		std::set<u64> leaves;
		size_t num_cids = 0;
		u64 nref = r.get<u64>();
		for(u64 i = 0; i < nref && r.ok; i++) {
			std::string uid;
			while(true) {
				int c = r.ch();
				while(c != EOF && isspace(c)) c = r.ch();
				if(c == EOF || c == '\0') break;
				uid.push_back((char)c);
			}
			if(uid.find("cid") == 0) num_cids++;
			u64 tid = r.get<u64>();
			ix.uid_to_tid.push_back(std::make_pair(uid, tid));
			leaves.insert(tid);
		}
		ix.compressed = num_cids >= 10;             // bt2_idx.h:661-663
		u64 ntid = r.get<u64>();
		while(ntid > 0 && r.ok) {
			u64 tid = r.get<u64>(); TaxNode n; n.parent = r.get<u64>(); n.rank = (uint8_t)r.get<uint16_t>();
			if(!r.ok) break;
			n.leaf = leaves.count(tid) ? 1 : 0;
			ix.tree[tid] = n;
			if(ix.tree.size() == ntid) break;
		}
		u64 nname = r.get<u64>();
		while(nname > 0 && r.ok) {
			u64 tid = r.get<u64>();
			if(!r.ok) break;
			std::string nm; int c = r.ch();
			while(c != EOF && isspace(c)) c = r.ch();
			while(c != EOF && !isspace(c)) { nm.push_back((char)c); c = r.ch(); }
			// the reference leaves the delimiter unread and then seekg(1)s over it: net effect = consumed
			std::replace(nm.begin(), nm.end(), '@', ' ');
			ix.name[tid] = nm;
			if(ix.name.size() == nname) break;
		}
		u64 nsize = r.get<u64>();
		while(nsize > 0 && r.ok) {
			u64 tid = r.get<u64>(); u64 sz = r.get<u64>();
			if(!r.ok) break;
			ix.size[tid] = sz;
			if(ix.size.size() == nsize) break;
		}
		// average genome size for internal ranks, bt2_idx.h:709-744
		uint8_t trn[RANK_MAX]; fill_tax_rank_num(trn);
		std::map<u64, u64> tid_count, new_size;
		for(std::map<u64, u64>::const_iterator it = ix.size.begin(); it != ix.size.end(); ++it) {
			u64 c_tid = it->first;
			if(!ix.tree.count(c_tid) || ix.tree[c_tid].parent == c_tid) continue;
			u64 add = it->second;
			const TaxNode& sn = ix.tree[c_tid];
			if(!((sn.rank == RANK_UNKNOWN && sn.leaf) || trn[sn.rank] < trn[RANK_SPECIES]) || sn.parent == c_tid) continue;
			c_tid = sn.parent;
			while(true) {
				std::map<u64, TaxNode>::const_iterator t = ix.tree.find(c_tid);
				if(t == ix.tree.end()) break;
				uint8_t rk = t->second.rank;
				if(rk == RANK_SPECIES || rk == RANK_GENUS || rk == RANK_FAMILY || rk == RANK_ORDER || rk == RANK_CLASS || rk == RANK_PHYLUM) {
					new_size[c_tid] += add; ++tid_count[c_tid];
				}
				if(c_tid == t->second.parent) break;
				c_tid = t->second.parent;
			}
		}
		for(std::map<u64, u64>::const_iterator it = tid_count.begin(); it != tid_count.end(); ++it)
			ix.size[it->first] = new_size[it->first] / it->second;
		build_paths(ix);
	}
	{   // bt2_idx.h:789-853; absent file => empty boundary set
		Reader r(base + ".4.cf");
		if(r.ok) {
			(void)r.get<u32>();
			u64 n = r.get<u64>();
			for(u64 i = 0; i < n && r.ok; i++) {
				u64 row = r.get<u64>(); u32 ref = r.get<u32>();
				if(!r.ok) break;
				ix.boundary[row] = ref;
				if(row > ix.lastBoundary) ix.lastBoundary = row;
			}
		}
	}
	return true;
}

// ---------------------------------------------------------------------------
// FM primitives
// ---------------------------------------------------------------------------
inline int bwt_char(const cfo_index& ix, u64 row) {      // rowL bt2_idx.h:2737
	u64 s = row / ix.sideBwtLen, off = row % ix.sideBwtLen;
	return (ix.ebwt[s * ix.sideSz + (off >> 2)] >> ((off & 3) * 2)) & 3;
}
// countBt2Side bt2_idx.h:2192-2227 + countUpTo :2364-2425 (result only; the word-wise
// popcount tricks of the reference are an implementation detail)
inline u64 lf(const cfo_index& ix, u64 row, int c) {
	u64 s = row / ix.sideBwtLen, off = row % ix.sideBwtLen;
	const uint8_t* side = &ix.ebwt[s * ix.sideSz];
	u64 n = 0;
	for(u64 i = 0; i < off; i++) n += (((side[i >> 2] >> ((i & 3) * 2)) & 3) == c);
	if(c == 0 && ix.zOff / ix.sideBwtLen == s && ix.zOff % ix.sideBwtLen < off) n--;  // '$' stored as A
	u64 occ; memcpy(&occ, side + ix.sideBwtSz + 8 * c, 8);
	return ix.fchr[c] + occ + n;
}
// faster word-based version used by the timed baseline; identical results (asserted in tests)
inline u64 lf_fast(const cfo_index& ix, u64 row, int c) {
	static const u64 ctab[4] = {0xffffffffffffffffull, 0xaaaaaaaaaaaaaaaaull, 0x5555555555555555ull, 0};
	u64 s = row / ix.sideBwtLen, off = row % ix.sideBwtLen;
	const uint8_t* side = &ix.ebwt[s * ix.sideSz];
	u64 n = 0, full = off >> 5, rem = off & 31;
	for(u64 w = 0; w < full; w++) {
		u64 dw; memcpy(&dw, side + 8 * w, 8);
		u64 x = dw ^ ctab[c];
		n += (u64)__builtin_popcountll((x >> 1) & x & 0x5555555555555555ull);
	}
	if(rem) {
		u64 dw; memcpy(&dw, side + 8 * full, 8);
		u64 x = dw ^ ctab[c];
		u64 m = (x >> 1) & x & 0x5555555555555555ull & (((u64)1 << (2 * rem)) - 1);
		n += (u64)__builtin_popcountll(m);
	}
	if(c == 0 && ix.zOff / ix.sideBwtLen == s && ix.zOff % ix.sideBwtLen < off) n--;
	u64 occ; memcpy(&occ, side + ix.sideBwtSz + 8 * c, 8);
	return ix.fchr[c] + occ + n;
}
inline u64 ftab_hi(const cfo_index& ix, u64 i) {         // bt2_idx.h:1878-1894
	u64 e = ix.ftab[i];
	return e <= ix.len ? e : ix.eftab[(e ^ OFF) * 2 + 1];
}
inline u64 ftab_lo(const cfo_index& ix, u64 i) {         // bt2_idx.h:1957-1973
	u64 e = ix.ftab[i];
	return e <= ix.len ? e : ix.eftab[(e ^ OFF) * 2];
}

struct Hit { u64 top, bot, bwoff, len; u64 size() const { return bot - top; } };
struct StrandHits {                                       // ReadBWTHit hi_aligner.h:150-319
	bool fw; u64 len, cur; bool done; std::vector<Hit> hits;
	void init(bool fw_, u64 len_) { fw = fw_; len = len_; cur = 0; done = false; hits.clear(); }
};

struct Ctx {
	const cfo_index& ix; const cfo_params& p; cfo_stats* st;
	std::set<u64> host, excluded;
	u64 ihits;
	Ctx(const cfo_index& ix_, const cfo_params& p_, cfo_stats* st_) : ix(ix_), p(p_), st(st_) {
		// ReportingParams aln_sink.h:580-588
		ihits = (u64)std::max(p.khits, 5) * (ix.compressed ? 4 : 40);
		// Classifier ctor classifier.h:157-201: every tree node with a listed id on its ancestor chain
		expand(p.host_taxids, p.n_host, host);
		expand(p.excluded_taxids, p.n_excluded, excluded);
	}
	void expand(const u64* ids, size_t n, std::set<u64>& out) {
		if(n == 0) return;
		for(std::map<u64, TaxNode>::const_iterator itr = ix.tree.begin(); itr != ix.tree.end(); ++itr) {
			u64 t = itr->first;
			while(true) {
				bool found = false;
				for(size_t k = 0; k < n; k++) if(t == ids[k]) { out.insert(itr->first); found = true; break; }
				if(found) break;
				std::map<u64, TaxNode>::const_iterator i2 = ix.tree.find(t);
				if(i2 == ix.tree.end()) break;
				if(t == i2->second.parent) break;
				t = i2->second.parent;
			}
		}
	}
};

// partialSearch hi_aligner.h:903-1031.  seq = strand sequence (patFw or patRc), 1 byte/base.
void partial_search(Ctx& cx, const uint8_t* seq, u64 len, StrandHits& H, bool ext = false) {
	const cfo_index& ix = cx.ix;
	const u64 ftabLen = (u64)ix.ftabChars;
	if(cx.st) { cx.st->partial_searches++; if(ext) cx.st->ext_searches++; }
	u64 offset = H.cur, dep = offset;
	u64 left = len - dep;
	if(left < ftabLen) {                                   // :939-949
		H.cur = H.len;
		Hit h = {OFF, OFF, (u32)offset, (u32)(H.cur - offset)}; H.hits.push_back(h);
		H.done = true; return;
	}
	for(u64 i = 0; i < ftabLen; i++) {                     // :951-966
		int c = seq[len - dep - 1 - i];
		if(c > 3) {
			H.cur += (i + 1);
			Hit h = {OFF, OFF, (u32)offset, (u32)(H.cur - offset)}; H.hits.push_back(h);
			if(H.cur >= H.len) H.done = true;
			return;
		}
	}
	// ftabLoHi bt2_idx.h:1931 with ftabSeqToInt :1830 (fw index, rev=false: leftmost base most significant)
	u64 fi = 0;
	for(u64 i = 0; i < ftabLen; i++) fi = (fi << 2) | seq[len - dep - ftabLen + i];
	u64 top = ftab_hi(ix, fi), bot = ftab_lo(ix, fi + 1);
	if(cx.st) cx.st->ftab_probes++;
	dep += ftabLen;
	if(bot <= top) {                                       // :971-982
		H.cur = dep;
		Hit h = {OFF, OFF, (u32)offset, (u32)(H.cur - offset)}; H.hits.push_back(h);
		if(H.cur >= H.len) H.done = true;
		return;
	}
	while(dep < len) {                                     // :985-1008
		int c = seq[len - dep - 1];
		u64 t = 0, b = 0;
		if(c <= 3) {
			if(bot - top != 1) {                           // bloc.valid(): HIER_INIT_LOCS :880
				t = lf_fast(ix, top, c); b = lf_fast(ix, bot, c);
				if(cx.st) {
					cx.st->lf_range_steps++;
					// initFromTopBot bt2_idx.h:326-349: bot reuses top's side iff charOff+spread < sideBwtLen
					bool same = (top % ix.sideBwtLen) + (bot - top) < ix.sideBwtLen;
					if(same) { cx.st->lf_range_same_side++; cx.st->sides_search += 1; } else cx.st->sides_search += 2;
				}
			} else {                                       // mapLF1 bt2_idx.h:2910
				if(cx.st) { cx.st->lf_single_steps++; cx.st->sides_search += 1; }
				if(bwt_char(ix, top) != c || top == ix.zOff) { t = b = 0; }
				else { t = lf_fast(ix, top, c); b = t + 1; }
			}
		}
		if(b <= t) break;
		top = t; bot = b; dep++;
	}
	Hit h = {top, bot, (u32)offset, (u32)(dep - offset)};  // :1011-1029 (bot > top always here)
	H.hits.push_back(h);
	H.cur = dep;
	if(H.cur >= H.len) H.done = true;
}

struct Mate { const uint8_t* fw; std::vector<uint8_t> rc; u64 len; StrandHits H[2]; };

// searchForwardAndReverse classifier.h:646-896
void search_fw_rc(Ctx& cx, Mate& m, u64 increment, int stop_before_trim = 0) {
	const u64 minHitLen = (u64)cx.p.min_hitlen, rdlen = m.len;
	bool done[2] = {false, false};
	size_t sum[2] = {0, 0};
	while(!done[0] || !done[1]) {
		for(u64 fwi = 0; fwi < 2; fwi++) {
			if(done[fwi]) continue;
			StrandHits& H = m.H[fwi];
			partial_search(cx, fwi == 0 ? m.fw : m.rc.data(), rdlen, H);
			Hit& last = H.hits.back();
			if(H.done) {
				done[fwi] = true;
				if(last.len >= minHitLen) sum[fwi] += last.len;
				continue;
			}
			if(last.len >= minHitLen) sum[fwi] += last.len;
			if(last.len > increment) H.cur = H.cur + 1;        // :727-761 (both branches +1)
			if(H.cur + minHitLen >= rdlen) { H.done = true; done[fwi] = true; continue; }
			if(last.len <= 3) --fwi;                           // unsigned wrap + loop ++ => repeat strand
		}
	}
	if(stop_before_trim == 2) return;
	// Extend partial hits :790-847
	if(sum[0] >= minHitLen && sum[1] >= minHitLen) {
		std::vector<Hit>& F = m.H[0].hits; std::vector<Hit>& R = m.H[1].hits;
		for(size_t i = 0; i < F.size(); i++) {
			Hit& hit = F[i];
			u64 len = hit.len, l = hit.bwoff, r = hit.bwoff + len;   // not refreshed after replacement
			for(size_t j = 0; j < R.size(); j++) {
				Hit& rchit = R[j];
				u64 rclen = rchit.len;
				if(len < minHitLen && rclen < minHitLen) continue;
				u64 rc_l = rdlen - rchit.bwoff - rchit.len, rc_r = rc_l + rclen;
				if(r <= rc_l) continue;
				if(rc_r <= l) continue;
				if(l == rc_l && r == rc_r) continue;
				if(l < rc_l && r > rc_r) continue;
				if(l > rc_l && r < rc_r) continue;
				if(l > rc_l) {
					StrandHits T; T.init(true, rdlen); T.cur = rc_l;
					partial_search(cx, m.fw, rdlen, T, true);
					const Hit& t = T.hits[0];
					if(t.len == len + l - rc_l) hit = t;
				}
				if(r > rc_r) {
					StrandHits T; T.init(false, rdlen); T.cur = rdlen - r;
					partial_search(cx, m.rc.data(), rdlen, T, true);
					const Hit& t = T.hits[0];
					if(t.len == rclen + r - rc_r) rchit = t;
				}
			}
		}
		// Remove twin hits mapped more than ihits times :850-870
		for(size_t i = 0; i < F.size(); i++) {
			Hit& hit = F[i];
			u64 len = hit.len, l = hit.bwoff, r = hit.bwoff + len;
			for(size_t j = 0; j < R.size(); j++) {
				Hit& rchit = R[j];
				u64 rclen = rchit.len;
				u64 rc_l = rdlen - rchit.bwoff - rchit.len, rc_r = rc_l + rclen;
				if(rc_l < l) break;
				if(len != rclen) continue;
				if(l == rc_l && r == rc_r && hit.size() + rchit.size() > cx.ihits) {
					Hit z = {0, 0, OFF, 0};                    // BWTHit::reset hi_aligner.h:63-72
					hit = z; rchit = z;
					break;
				}
			}
		}
	}
	if(stop_before_trim == 1) return;
	// Trim partial hits :874-895
	for(int fwi = 0; fwi < 2; fwi++) {
		std::vector<Hit>& L = m.H[fwi].hits;
		if(L.size() < 2) continue;
		for(size_t i = 0; i + 1 < L.size(); i++) {
			Hit& a = L[i];
			for(size_t j = i + 1; j < L.size(); j++) {
				Hit& b = L[j];
				if(a.bwoff >= b.bwoff) { a.len = 0; break; }
				if(a.bwoff + a.len <= b.bwoff) break;
				if(a.len >= b.len) { u64 e = b.bwoff + b.len; b.bwoff = a.bwoff + a.len; b.len = e - b.bwoff; }
				else a.len = b.bwoff - a.bwoff;
			}
		}
	}
}

// getForwardOrReverseHit classifier.h:898-941
std::pair<int, int> choose_strand(const Mate& m, u64 minHitLen) {
	u64 avg[2] = {0, 0}, mx[2] = {0, 0};
	for(int fwi = 0; fwi < 2; fwi++) {
		u64 nh = 0, tot = 0;
		for(size_t i = 0; i < m.H[fwi].hits.size(); i++) {
			u64 len = m.H[fwi].hits[i].len;
			if(len < minHitLen) continue;
			tot += (len - 15) * (len - 15);
			if(len > mx[fwi]) mx[fwi] = len;
			nh++;
		}
		if(nh > 0) avg[fwi] = tot;
	}
	if(avg[0] != avg[1]) { int f = avg[0] > avg[1] ? 0 : 1; return std::make_pair(f, f + 1); }
	if(mx[0] != mx[1])   { int f = mx[0] > mx[1] ? 0 : 1;   return std::make_pair(f, f + 1); }
	return std::make_pair(0, 2);
}

struct CmpHits {                                          // compareBWTHits classifier.h:1058-1086
	bool operator()(const Hit& a, const Hit& b) const {
		if(a.len >= 22 || b.len >= 22) {
			if(a.len >= 22 && b.len >= 22) {
				if(a.size() < b.size()) return true;
				if(a.size() > b.size()) return false;
			}
			if(b.len < a.len) return true;
			if(b.len > a.len) return false;
		}
		if(b.len * a.size() < a.len * b.size()) return true;
		if(b.len * a.size() > a.len * b.size()) return false;
		if(a.size() < b.size()) return true;
		if(a.size() > b.size()) return false;
		if(b.len < a.len) return true;
		if(b.len > a.len) return false;
		return false;
	}
};

// resolve one SA row: GWState::init/advance group_walk.h:475-705,855-1016 restated per row,
// tryOffset bt2_idx.h:1980-2014; no "+steps" under -DCENTRIFUGE (group_walk.h:508-512)
u64 resolve_row(const cfo_index& ix, u64 row, u64* steps) {
	while(true) {
		if(row == ix.zOff) return 0;
		if((row & ix.offMask) == row) {
			u64 i = row >> ix.offRate;
			return ix.offw ? (u64)ix.offs32[i] : (u64)ix.offs16[i];
		}
		if(ix.lastBoundary > 0 && row <= ix.lastBoundary) {
			std::map<u64, u32>::const_iterator it = ix.boundary.find(row);
			if(it != ix.boundary.end()) return ix.offw ? (u64)it->second : (u64)(uint16_t)it->second;
		}
		row = lf_fast(ix, row, bwt_char(ix, row));
		if(steps) (*steps)++;
	}
}

struct HitCount {                                         // classifier.h:31-121
	u64 uniqueID, taxID; u32 count, score; u32 scores[2][2]; double summedHitLen; double lens[2][2];
	u32 timeStamp; bool leaf; u32 num_leaves; uint8_t rank; std::vector<u64> path;
};

struct Unit { Mate m[2]; int nm; };

// Classifier::go classifier.h:212-571.  Returns records in hitMap order.
void classify_unit(Ctx& cx, Unit& u, std::vector<cfo_rec>& out) {
	const cfo_index& ix = cx.ix;
	const u64 minHitLen = (u64)cx.p.min_hitlen, khits = (u64)cx.p.khits;
	std::vector<HitCount> hitMap;
	const u64 increment = (2 * minHitLen <= 33) ? 10 : (2 * minHitLen - 33);
	u64 maxG = khits;
	u32 ts = 0;
	const bool paired = u.nm == 2;
	for(int rdi = 0; rdi < u.nm; rdi++) {
		Mate& m = u.m[rdi];
		search_fw_rc(cx, m, increment);
		std::pair<int, int> fwp = choose_strand(m, minHitLen);
		for(int fwi = fwp.first; fwi < fwp.second; fwi++) {
			std::vector<Hit>& L = m.H[fwi].hits;
			size_t n = L.size();
			for(size_t hi = 0; hi < n; hi++)
				if(L[hi].len >= minHitLen && L[hi].size() > maxG) maxG = L[hi].size();
			if(maxG > khits) maxG += khits;
			std::sort(L.begin(), L.end(), CmpHits());          // EList::sort ds.h:775-778
			size_t genomeHitCnt = 0;
			for(size_t hi = 0; hi < n; hi++, ts++) {
				const Hit& ph = L[hi];
				if(ph.len <= minHitLen) continue;
				if(ph.size() == 0) continue;
				u64 nelt = std::min<u64>(ph.size(), maxG);      // getGenomeIdx :592-593
				if(cx.st) cx.st->hits_resolved++;
				std::vector<u64> ids(nelt);
				for(u64 e = 0; e < nelt; e++) {
					u64 steps = 0;
					ids[e] = resolve_row(ix, ph.top + e, &steps);
					if(cx.st) { cx.st->walk_steps += steps; cx.st->rows_resolved++; }
				}
				if(nelt > cx.ihits) continue;                   // :299
				std::vector<std::pair<u64, u64> > cid;
				for(u64 k = 0; k < nelt; k++, genomeHitCnt++) {
					u64 ref = ids[k];
					u64 taxID = ref < ix.uid_to_tid.size() ? ix.uid_to_tid[ref].second : 0;
					bool found = false;
					for(size_t k2 = 0; k2 < cid.size(); k2++) if(cid[k2].first == ref) { found = true; break; }
					if(found) continue;
					cid.push_back(std::make_pair(ref, taxID));
				}
				u32 hitScore = (u32)((ph.len - 15) * (ph.len - 15));
				double wl = (double)ph.len;
				for(size_t k = 0; k < cid.size(); k++) {
					u64 uniqueID = cid[k].first, taxID = cid[k].second;
					if(cx.excluded.count(taxID)) continue;
					// addHitToHitMap classifier.h:982-1050
					std::vector<u64> path;
					std::map<u64, u32>::const_iterator pit = ix.tid_to_pid.find(taxID);
					if(pit != ix.tid_to_pid.end()) path = ix.paths[pit->second];
					uint8_t rank = (uint8_t)cx.p.class_rank_slot;
					if(rank > 0) {
						for(; rank < path.size(); rank++) if(path[rank] != 0) { taxID = path[rank]; break; }
					}
					size_t idx = 0;
					for(; idx < hitMap.size(); ++idx) {
						bool same = (rank == 0) ? (uniqueID == hitMap[idx].uniqueID) : (taxID == hitMap[idx].taxID);
						if(same) {
							if(hitMap[idx].timeStamp != ts) {
								hitMap[idx].count += 1;
								hitMap[idx].scores[rdi][fwi] += hitScore;
								hitMap[idx].lens[rdi][fwi] += wl;
								hitMap[idx].timeStamp = ts;
							}
							break;
						}
					}
					if(idx >= hitMap.size()) {
						HitCount hc; memset(hc.scores, 0, sizeof(hc.scores)); memset(hc.lens, 0, sizeof(hc.lens));
						hc.score = 0; hc.summedHitLen = 0.0; hc.leaf = true; hc.num_leaves = 1;
						hc.uniqueID = uniqueID; hc.count = 1; hc.scores[rdi][fwi] = hitScore; hc.lens[rdi][fwi] = wl;
						hc.timeStamp = ts; hc.path = path; hc.rank = rank; hc.taxID = taxID;
						hitMap.push_back(hc);
					}
				}
				if(genomeHitCnt >= maxG) break;
			}
		}
	}
	for(size_t i = 0; i < hitMap.size(); i++) {               // HitCount::finalize :86-120
		HitCount& h = hitMap[i];
		if(paired) {
			h.score = std::max(h.scores[0][0], h.scores[0][1]) + std::max(h.scores[1][0], h.scores[1][1]);
			h.summedHitLen = std::max(h.lens[0][0], h.lens[0][1]) + std::max(h.lens[1][0], h.lens[1][1]);
		} else {
			h.score = std::max(h.scores[0][0], h.scores[0][1]);
			h.summedHitLen = std::max(h.lens[0][0], h.lens[0][1]);
		}
	}
	int64_t best_score = 0; bool only_host = false;           // :385-394
	for(size_t gi = 0; gi < hitMap.size(); gi++) {
		if((int64_t)hitMap[gi].score > best_score) { best_score = hitMap[gi].score; only_host = cx.host.count(hitMap[gi].taxID) > 0; }
		else if((int64_t)hitMap[gi].score == best_score) only_host |= cx.host.count(hitMap[gi].taxID) > 0;
	}
	bool unclassified = false;
	if(!only_host && hitMap.size() > khits) {                 // :399-515
		u32 best = hitMap[0].score;
		for(size_t i = 1; i < hitMap.size(); i++) if(best < hitMap[i].score) best = hitMap[i].score;
		for(int i = 0; i < (int)hitMap.size(); i++) {
			if(hitMap[i].score < best) {
				if(i + 1 < (int)hitMap.size()) hitMap[i] = hitMap.back();
				hitMap.pop_back(); i--;
			}
		}
		if(!cx.p.tree_traverse && hitMap.size() > khits) unclassified = true;
		uint8_t rank = 0;
		std::vector<std::pair<u32, u64> > tc;
		while(!unclassified && hitMap.size() > khits) {
			tc.clear();
			for(size_t i = 0; i < hitMap.size(); i++) {
				HitCount& h = hitMap[i];
				while(h.rank < rank) {
					if((size_t)h.rank + 1 >= h.path.size()) { h.rank = 255; break; }
					h.rank += 1; h.taxID = h.path[h.rank]; h.leaf = false;
				}
				if(h.rank > rank) continue;
				u64 parent = ((size_t)rank + 1 >= h.path.size()) ? 1 : h.path[rank + 1];
				if(parent == 0) continue;
				size_t j = 0;
				for(; j < tc.size(); j++) if(tc[j].second == parent) { tc[j].first += 1; break; }
				if(j == tc.size()) tc.push_back(std::make_pair((u32)1, parent));
			}
			if(tc.empty()) {
				if(rank < hitMap[0].path.size()) { rank++; continue; } else break;
			}
			std::sort(tc.begin(), tc.end());
			size_t j = tc.size();
			while(j-- > 0) {
				u64 parent = tc[j].second;
				for(size_t i = 0; i < hitMap.size(); i++) {
					HitCount& h = hitMap[i];
					if(h.rank != rank) continue;
					u64 cur_parent = ((size_t)rank + 1 >= h.path.size()) ? 1 : h.path[rank + 1];
					if(parent == cur_parent) { h.uniqueID = OFF; h.rank = rank + 1; h.taxID = parent; h.leaf = false; }
				}
				bool first = true; size_t rep_i = hitMap.size();
				for(size_t i = 0; i < hitMap.size(); i++) {
					if(parent == hitMap[i].taxID) {
						if(!first) {
							hitMap[rep_i].num_leaves += hitMap[i].num_leaves;
							if(i + 1 < hitMap.size()) hitMap[i] = hitMap.back();
							hitMap.pop_back(); i--;
						} else { first = false; rep_i = i; }
					}
				}
				if(hitMap.size() <= khits) break;
			}
			++rank;
			if(rank > hitMap[0].path.size()) break;
		}
	}
	if(!only_host && hitMap.size() > khits) unclassified = true;   // :516-520
	if(unclassified) return;
	for(size_t gi = 0; gi < hitMap.size(); gi++) {                 // :538-565
		const HitCount& h = hitMap[gi];
		if(only_host && !cx.host.count(h.taxID)) continue;
		cfo_rec r; r.taxid = h.taxID; r.score = h.score; r.hitlen = (u32)(u64)h.summedHitLen;
		r.uid = h.uniqueID < ix.uid_to_tid.size() ? (u32)h.uniqueID : 0xFFFFFFFFu; r.pad = 0;
		out.push_back(r);
	}
}

void make_rc(const uint8_t* fw, u64 len, std::vector<uint8_t>& rc) {  // Read::constructRevComps read.h
	rc.resize(len);
	for(u64 i = 0; i < len; i++) { uint8_t c = fw[len - 1 - i]; rc[i] = c > 3 ? 4 : (uint8_t)(3 - c); }
}

void setup_unit(Unit& u, const uint8_t* b1, u32 l1, const uint8_t* b2, u32 l2, uint8_t flags) {
	// centrifuge.cpp:2678-2690: both pass -> initReads; one passes -> initRead on that mate
	u.nm = 0;
	if(flags & 1) { Mate& m = u.m[u.nm++]; m.fw = b1; m.len = l1; make_rc(b1, l1, m.rc); }
	if((flags & 2) && b2) { Mate& m = u.m[u.nm++]; m.fw = b2; m.len = l2; make_rc(b2, l2, m.rc); }
	for(int i = 0; i < u.nm; i++) { u.m[i].H[0].init(true, u.m[i].len); u.m[i].H[1].init(false, u.m[i].len); }
}

}  // namespace

// ---------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------
extern "C" cfo_index* cfo_index_load(const char* basename, char* err, size_t errlen) {
	cfo_index* ix = new cfo_index();
	std::string e;
	if(!load_index(*ix, basename, e)) {
		if(err && errlen) { strncpy(err, e.c_str(), errlen - 1); err[errlen - 1] = 0; }
		delete ix; return NULL;
	}
	return ix;
}
extern "C" void cfo_index_free(cfo_index* ix) { delete ix; }
extern "C" uint64_t cfo_index_len(const cfo_index* ix) { return ix->len; }
extern "C" uint64_t cfo_index_nseq(const cfo_index* ix) { return ix->uid_to_tid.size(); }
extern "C" int cfo_index_sample_width(const cfo_index* ix) { return ix->offw ? 4 : 2; }
extern "C" int cfo_index_compressed(const cfo_index* ix) { return ix->compressed ? 1 : 0; }
extern "C" uint64_t cfo_lf(const cfo_index* ix, uint64_t row, int c) { return lf(*ix, row, c); }
extern "C" int cfo_bwt_char(const cfo_index* ix, uint64_t row) { return bwt_char(*ix, row); }
extern "C" uint64_t cfo_resolve(const cfo_index* ix, uint64_t row, uint64_t* steps) { return resolve_row(*ix, row, steps); }
extern "C" void cfo_ftab_lohi(const cfo_index* ix, const uint8_t* s, uint64_t* top, uint64_t* bot) {
	u64 fi = 0; for(int i = 0; i < ix->ftabChars; i++) fi = (fi << 2) | s[i];
	*top = ftab_hi(*ix, fi); *bot = ftab_lo(*ix, fi + 1);
}

extern "C" int64_t cfo_classify(const cfo_index* ix, const cfo_params* p,
                                const uint8_t* bases, const uint64_t* off1, const uint32_t* len1,
                                const uint64_t* off2, const uint32_t* len2, const uint8_t* flags,
                                size_t n, uint32_t* out_n, cfo_rec* out, size_t cap, cfo_stats* stats) {
	Ctx cx(*ix, *p, stats);
	size_t total = 0;
	std::vector<cfo_rec> recs;
	for(size_t i = 0; i < n; i++) {
		recs.clear();
		uint8_t fl = flags ? flags[i] : 1;
		u32 l2 = (off2 && len2) ? len2[i] : 0;
		bool has2 = off2 != NULL && len2 != NULL && (fl & 4);   // bit2: unit is a pair
		if(!has2) fl &= 1;
		if(fl & 3) {
			Unit u;
			setup_unit(u, bases + off1[i], len1[i], has2 ? bases + off2[i] : NULL, l2, fl);
			classify_unit(cx, u, recs);
			if(stats) stats->reads++;
		}
		out_n[i] = (u32)recs.size();
		if(total + recs.size() > cap) return -1;
		for(size_t k = 0; k < recs.size(); k++) out[total++] = recs[k];
	}
	return (int64_t)total;
}

extern "C" int cfo_search_dump(const cfo_index* ix, const cfo_params* p, const uint8_t* bases, uint32_t len,
                               int after_trim, uint32_t n_hits[2], uint64_t* top, uint64_t* bot,
                               uint32_t* bwoff, uint32_t* hlen, size_t cap) {
	Ctx cx(*ix, *p, NULL);
	Unit u; setup_unit(u, bases, len, NULL, 0, 1);
	const u64 minHitLen = (u64)p->min_hitlen;
	const u64 increment = (2 * minHitLen <= 33) ? 10 : (2 * minHitLen - 33);
	search_fw_rc(cx, u.m[0], increment, after_trim ? 0 : 2);
	for(int s = 0; s < 2; s++) {
		const std::vector<Hit>& L = u.m[0].H[s].hits;
		n_hits[s] = (u32)L.size();
		for(size_t i = 0; i < L.size() && i < cap; i++) {
			top[s * cap + i] = L[i].top; bot[s * cap + i] = L[i].bot;
			bwoff[s * cap + i] = (u32)L[i].bwoff; hlen[s * cap + i] = (u32)L[i].len;
		}
	}
	return 0;
}

// ---------------------------------------------------------------------------
// File driver: read parsing (pat.cpp:725-849 FASTA, :852-1157 FASTQ), per-read seed
// (pat.h:55-91), N filter (scoring.cpp:104-168, NCEIL=L,0,0.15 aligner_seed_policy.cpp:296),
// selectByScore (aln_sink.h:1861-1927), TSV (aln_sink.h:2202-2337), report
// (centrifuge.cpp:3231-3319), abundance EM (aln_sink.h:196-495).
// ---------------------------------------------------------------------------
namespace {

struct RawRead { std::string name; std::vector<uint8_t> seq; std::string qual; };

static const uint8_t* asc2dna_tab() {                    // alphabet.cpp:298-319
	static uint8_t t[256]; static bool init = false;
	if(!init) { memset(t, 0, 256); t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; t['N'] = t['n'] = 4; init = true; }
	return t;
}
static bool dnacat(int c) {                              // asc2dnacat > 0, alphabet.cpp:36-58
	if(c == '-') return true;                              // category 3, mapped to A by asc2dna
	int u = toupper(c);
	return c > 0 && u != 0 && strchr("ABCDGHKMNRSTVWXY", u) != NULL;
}

struct FileBuf {
	FILE* f; int pk; bool havepk;
	explicit FileBuf(const char* p) : f(fopen(p, "rb")), pk(0), havepk(false) {}
	~FileBuf() { if(f) fclose(f); }
	int get() { if(havepk) { havepk = false; return pk; } return fgetc(f); }
	int peek() { if(!havepk) { pk = fgetc(f); havepk = true; } return pk; }
};

// returns false at EOF
bool read_fasta(FileBuf& fb, RawRead& r, u64 readCnt, bool& first, bool& empty, int trim5 = 0, int trim3 = 0) {
	const uint8_t* a2d = asc2dna_tab();
	r.name.clear(); r.seq.clear(); r.qual.clear(); empty = false;
	int c = fb.get();
	if(c < 0) return false;
	while(c == '#' || c == ';' || c == '\r' || c == '\n') {
		// FileBuf::peekUptoNewline (filebuf.h:306-318): drop the rest of the line the cursor is in, then every line end.
		// For a comment that is the comment line; after a *line end* (only possible before the first record) it is
		// the whole next line -- a leading blank line makes the reference swallow the first header (pat.cpp:744-748).
		while(true) { int d = fb.peek(); if(d < 0 || d == '\n' || d == '\r') break; fb.get(); }
		while(fb.peek() == '\n' || fb.peek() == '\r') fb.get();
		c = fb.get();
		if(c < 0 && !first) return false;
		if(c < 0) break;
	}
	if(first) { if(c != '>') { fprintf(stderr, "Error: reads file does not look like a FASTA file\n"); exit(1); } first = false; }
	c = fb.get();
	while(true) {
		if(c < 0) return false;
		if(c == '\n' || c == '\r') {
			while(c == '\n' || c == '\r') {
				if(fb.peek() == '>') break;
				c = fb.get();
				if(c < 0) return false;
			}
			break;
		}
		r.name.push_back((char)c);
		if(fb.peek() == '>') break;
		c = fb.get();
	}
	if(c == '>') { empty = true; return true; }            // never true here in practice (c is not advanced onto '>')
	if(fb.peek() == '>' && (c == '\n' || c == '\r')) { empty = true; return true; }
	int begin = 0;
	while(c != '>' && c >= 0) {
		if(dnacat(c) && begin++ >= trim5) { r.seq.push_back(a2d[c]); r.qual.push_back('I'); }   // pat.cpp:826-830
		if(fb.peek() == '>') break;
		c = fb.get();
	}
	if(trim3 > 0) { const size_t keep = r.seq.size() > (size_t)trim3 ? r.seq.size() - trim3 : 0; r.seq.resize(keep); r.qual.resize(keep); }   // trimEnd :833-834
	if(r.name.empty()) { char b[32]; snprintf(b, sizeof b, "%llu", (unsigned long long)readCnt); r.name = b; }
	return true;
}

bool read_fastq(FileBuf& fb, RawRead& r, u64 readCnt, bool& first, int trim5 = 0, int trim3 = 0) {
	const uint8_t* a2d = asc2dna_tab();
	r.name.clear(); r.seq.clear(); r.qual.clear();
	int c;
	if(first) {
		c = fb.get();
		while(c == '\n' || c == '\r') c = fb.get();
		if(c < 0) return false;
		if(c != '@') { fprintf(stderr, "Error: reads file does not look like a FASTQ file\n"); exit(1); }
		first = false;
	}
	while(true) {
		c = fb.get();
		if(c < 0) return false;
		if(c == '\n' || c == '\r') {
			while(c == '\n' || c == '\r') { c = fb.get(); if(c < 0) return false; }
			break;
		}
		r.name.push_back((char)c);
	}
	int nread = 0;
	while(c != '+') {
		if(c == '.') c = 'N';
		if(isalpha(c)) { if(nread >= trim5) r.seq.push_back(a2d[c]); nread++; }      // pat.cpp:939-946
		c = fb.get();
		if(c < 0) return false;
	}
	if(trim3 > 0) r.seq.resize(r.seq.size() > (size_t)trim3 ? r.seq.size() - trim3 : 0);   // :970-982
	while(true) { int d = fb.get(); if(d < 0 || d == '\n' || d == '\r') { while(fb.peek() == '\n' || fb.peek() == '\r') fb.get(); break; } }
	if(nread == 0) { if(fb.peek() == '@') fb.get(); return true; }
	int qi = 0;
	while(true) {                                          // pat.cpp:1042-1078, phred33 (qual.h:136-142)
		c = fb.get();
		if(c == ' ') {
			fprintf(stderr, "Error: Encountered one or more spaces while parsing the quality string for read %s.  If this is a FASTQ file with integer (non-ASCII-encoded) qualities, try re-running with the --integer-quals option.\n", r.name.c_str());
			exit(1);
		}
		if(c < 0) break;
		if(c != '\r' && c != '\n') {
			if(qi >= trim5) {
				if((int)(signed char)c < 33) { fprintf(stderr, "Saw ASCII character %d but expected 33-based Phred qual.\n", (int)(signed char)c); exit(1); }
				r.qual.push_back((char)c);
			}
			qi++;
		} else break;
	}
	if(trim3 > 0) r.qual.resize(r.qual.size() > (size_t)trim3 ? r.qual.size() - trim3 : 0);
	if(r.qual.size() < r.seq.size()) { fprintf(stderr, "Error: Read %s has more read characters than quality values.\n", r.name.c_str()); exit(1); }
	if(r.qual.size() > r.seq.size() + 1) { fprintf(stderr, "Error: Read %s has more quality values than read characters.\n", r.name.c_str()); exit(1); }
	if(r.qual.size() > r.seq.size()) r.qual.resize(r.seq.size());
	while(fb.peek() == '\n' || fb.peek() == '\r') fb.get();
	c = fb.get();                                          // '@' of the next record or EOF
	if(r.name.empty()) { char b[32]; snprintf(b, sizeof b, "%llu", (unsigned long long)readCnt); r.name = b; }
	return true;
}

u32 gen_rand_seed(const RawRead& r, u32 seed) {            // pat.h:55-91
	u32 rseed = (seed + 101) * 59 * 61 * 67 * 71 * 73 * 79 * 83;
	size_t qlen = r.seq.size();
	for(size_t i = 0; i < qlen; i++) { int p = (int)r.seq[i]; rseed ^= ((u32)p << ((i & 15) << 1)); }
	for(size_t i = 0; i < qlen; i++) { int p = (int)(unsigned char)r.qual[i]; rseed ^= ((u32)p << ((i & 3) << 3)); }
	for(size_t i = 0; i < r.name.size(); i++) {
		int p = (int)r.name[i];
		if(p == '/') break;
		rseed ^= ((u32)p << ((i & 3) << 3));
	}
	return rseed;
}

struct Rng {                                               // random_source.h:34-61
	u32 last;
	void init(u32 s) { last = s; }
	u32 next() { u32 ret; last = 1664525u * last + 1013904223u; ret = last >> 16; last = 1664525u * last + 1013904223u; ret ^= last; return ret; }
};

bool n_filter(const std::vector<uint8_t>& s) {             // Scoring::nFilter scoring.cpp:104-117
	size_t maxns = (size_t)std::max(0.0, std::min(std::numeric_limits<double>::max(), 0.0 + 0.15 * (double)s.size()));
	size_t ns = 0;
	for(size_t i = 0; i < s.size(); i++) if(s[i] == 4) { ns++; if(ns > maxns) return false; }
	return true;
}

struct ReadCounts { u32 n_reads, n_unique_reads; };
struct IDs { std::vector<u64> ids;
	bool operator<(const IDs& o) const {                   // aln_sink.h:63-71
		if(ids.size() != o.ids.size()) return ids.size() < o.ids.size();
		for(size_t i = 0; i < ids.size(); i++) if(ids[i] != o.ids[i]) return ids[i] < o.ids[i];
		return false;
	}
};
struct Species {
	std::map<u64, ReadCounts> counts; std::map<IDs, u64> observed; IDs cur;
	std::map<u64, double> abundance, abundance_len;
	void add(u64 taxID, int64_t score, int64_t max_score, u32 nresult) {   // aln_sink.h:142-172
		ReadCounts& rc = counts[taxID]; rc.n_reads += 1; if(nresult == 1) rc.n_unique_reads += 1;
		if(score >= max_score) {
			cur.ids.push_back(taxID);
			if(cur.ids.size() == nresult) { std::sort(cur.ids.begin(), cur.ids.end()); observed[cur] += 1; cur.ids.clear(); }
		}
	}
};

void em_step(const std::map<IDs, u64>& observed, const std::map<u64, std::vector<u64> >& anc,
             const std::map<u64, u64>& t2n, const std::vector<double>& p, std::vector<double>& pn,
             const std::vector<size_t>& len) {           // aln_sink.h:196-272
	std::fill(pn.begin(), pn.end(), 0.0);
	for(std::map<IDs, u64>::const_iterator itr = observed.begin(); itr != observed.end(); ++itr) {
		const std::vector<u64>& ids = itr->first.ids; u64 count = itr->second; double psum = 0.0;
		for(size_t i = 0; i < ids.size(); i++) {
			std::map<u64, u64>::const_iterator id = t2n.find(ids[i]);
			if(id != t2n.end()) psum += p[id->second];
			else {
				std::map<u64, std::vector<u64> >::const_iterator a = anc.find(ids[i]);
				if(a == anc.end()) continue;
				for(size_t c = 0; c < a->second.size(); c++) { std::map<u64, u64>::const_iterator ci = t2n.find(a->second[c]); if(ci == t2n.end()) continue; psum += p[ci->second]; }
			}
		}
		if(psum == 0.0) continue;
		for(size_t i = 0; i < ids.size(); i++) {
			std::map<u64, u64>::const_iterator id = t2n.find(ids[i]);
			if(id != t2n.end()) pn[id->second] += (count * (p[id->second] / psum));
			else {
				std::map<u64, std::vector<u64> >::const_iterator a = anc.find(ids[i]);
				if(a == anc.end()) continue;
				for(size_t c = 0; c < a->second.size(); c++) { std::map<u64, u64>::const_iterator ci = t2n.find(a->second[c]); if(ci == t2n.end()) continue; pn[ci->second] += (count * (p[ci->second] / psum)); }
			}
		}
	}
	double sum = 0.0;
	for(size_t i = 0; i < pn.size(); i++) sum += (pn[i] / len[i]);
	for(size_t i = 0; i < pn.size(); i++) pn[i] = pn[i] / len[i] / sum;
}

void calc_abundance(const cfo_index& ix, Species& sp) {     // aln_sink.h:274-495
	const std::map<u64, TaxNode>& tree = ix.tree;
	std::set<u64> leaves;
	for(std::map<IDs, u64>::iterator itr = sp.observed.begin(); itr != sp.observed.end(); ++itr)
		for(size_t i = 0; i < itr->first.ids.size(); i++) {
			std::map<u64, TaxNode>::const_iterator t = tree.find(itr->first.ids[i]);
			if(t == tree.end()) continue;
			if(!t->second.leaf) continue;
			leaves.insert(t->first);
		}
	std::map<u64, std::vector<u64> > anc;
	for(std::map<IDs, u64>::iterator itr = sp.observed.begin(); itr != sp.observed.end(); ++itr)
		for(size_t i = 0; i < itr->first.ids.size(); i++) {
			u64 tid = itr->first.ids[i];
			if(leaves.count(tid)) continue;
			if(anc.count(tid)) continue;
			anc[tid].clear();
			for(std::set<u64>::const_iterator l = leaves.begin(); l != leaves.end(); ++l) {
				u64 t2 = *l, tmp = t2;
				while(true) {
					std::map<u64, TaxNode>::const_iterator t = tree.find(tmp);
					if(t == tree.end()) break;
					if(tid == t->second.parent) anc[tid].push_back(t2);
					if(tmp == t->second.parent) break;
					tmp = t->second.parent;
				}
			}
			std::sort(anc[tid].begin(), anc[tid].end());
		}
	std::map<u64, u64> t2n; std::vector<double> p; std::vector<size_t> len;
	for(std::map<IDs, u64>::iterator itr = sp.observed.begin(); itr != sp.observed.end(); ++itr) {
		const std::vector<u64>& ids = itr->first.ids; u64 count = itr->second;
		for(size_t i = 0; i < ids.size(); i++) {
			u64 tid = ids[i];
			if(!leaves.count(tid)) continue;
			if(!t2n.count(tid)) {
				t2n[tid] = p.size();
				p.push_back(1.0 / ids.size() * count);
				std::map<u64, u64>::const_iterator s = ix.size.find(tid);
				len.push_back(s != ix.size.end() ? (size_t)s->second : std::numeric_limits<size_t>::max());
			} else p[t2n[tid]] += (1.0 / ids.size() * count);
		}
	}
	{ double sum = 0.0; for(size_t i = 0; i < p.size(); i++) sum += (p[i] / len[i]); for(size_t i = 0; i < p.size(); i++) p[i] = (p[i] / len[i]) / sum; }
	std::vector<double> pn(p.size()), pn2(p.size()), pr(p.size()), pv(p.size());
	size_t it = 0; double diff = 0.0;
	while(true) {
		em_step(sp.observed, anc, t2n, p, pn, len);
		em_step(sp.observed, anc, t2n, pn, pn2, len);
		double ssr = 0.0, ssv = 0.0;
		for(size_t i = 0; i < p.size(); i++) { pr[i] = pn[i] - p[i]; ssr += pr[i] * pr[i]; pv[i] = pn2[i] - pn[i] - pr[i]; ssv += pv[i] * pv[i]; }
		if(ssv > 0.0) {
			double g = -sqrt(ssr / ssv);
			for(size_t i = 0; i < p.size(); i++) pn2[i] = std::max(0.0, p[i] - 2 * g * pr[i] + g * g * pv[i]);
			em_step(sp.observed, anc, t2n, pn2, pn, len);
		}
		diff = 0.0;
		for(size_t i = 0; i < p.size(); i++) diff += (p[i] > pn[i] ? p[i] - pn[i] : pn[i] - p[i]);
		if(diff < 0.0000000001) break;
		if(++it >= 10000) break;
		p = pn;
	}
	sp.abundance_len.clear(); sp.abundance.clear();
	double sum = 0.0;
	for(std::map<u64, u64>::iterator i = t2n.begin(); i != t2n.end(); ++i) { sp.abundance_len[i->first] = p[i->second]; sum += p[i->second] * len[i->second]; }
	for(std::map<u64, u64>::iterator i = t2n.begin(); i != t2n.end(); ++i) sp.abundance[i->first] = (p[i->second] * len[i->second]) / sum;
}

void append_read_id(std::string& o, const std::string& nm) {   // aln_sink.h:2202-2217
	size_t n = nm.size();
	if(n >= 2 && nm[n - 2] == '/' && (nm[n - 1] == '1' || nm[n - 1] == '2' || nm[n - 1] == '3')) n -= 2;
	for(size_t i = 0; i < n; i++) { if(isspace((unsigned char)nm[i])) break; o.push_back(nm[i]); }
}
void append_taxid(std::string& o, u64 tid) {                   // aln_sink.h:2237-2250
	char b[64]; u64 t1 = tid & 0xffffffffull, t2 = tid >> 32;
	snprintf(b, sizeof b, "%llu", (unsigned long long)t1); o += b;
	if(t2 > 0) { snprintf(b, sizeof b, ".%llu", (unsigned long long)t2); o += b; }
}

std::vector<u64> parse_ids(const char* s) {
	std::vector<u64> v; std::string t(s); std::stringstream ss(t); std::string tok;
	while(std::getline(ss, tok, ',')) if(!tok.empty()) v.push_back(strtoull(tok.c_str(), NULL, 10));
	return v;
}

}  // namespace

extern "C" int cfo_main(int argc, const char** argv) {
	std::string idx, u, m1, m2, out = "-", report = "centrifuge_report.tsv", statsf, dumpf;
	bool fasta = false, abundance = true; int trim5 = 0, trim3 = 0;
	cfo_params p; memset(&p, 0, sizeof p); p.khits = 5; p.min_hitlen = 22; p.tree_traverse = 1; p.class_rank_slot = 0;
	std::vector<u64> host, excl;
	for(int i = 1; i < argc; i++) {
		std::string a = argv[i];
		#define NEXT (i + 1 < argc ? argv[++i] : "")
		if(a == "-x") idx = NEXT; else if(a == "-U") u = NEXT; else if(a == "-1") m1 = NEXT; else if(a == "-2") m2 = NEXT;
		else if(a == "-f") fasta = true; else if(a == "-q") fasta = false; else if(a == "-S") out = NEXT;
		else if(a == "--report-file") report = NEXT; else if(a == "-k") p.khits = atoi(NEXT);
		else if(a == "--min-hitlen") p.min_hitlen = atoi(NEXT); else if(a == "--no-traverse") p.tree_traverse = 0;
		else if(a == "--host-taxids") host = parse_ids(NEXT); else if(a == "--exclude-taxids") excl = parse_ids(NEXT);
		else if(a == "--classification-rank") { uint8_t r = rank_to_pathID(rank_id(NEXT)); p.class_rank_slot = r; }
		else if(a == "--no-abundance") abundance = false; else if(a == "-p") (void)NEXT;
		else if(a == "--stats") statsf = NEXT;
		else if(a == "-5" || a == "--trim5") trim5 = atoi(NEXT); else if(a == "-3" || a == "--trim3") trim3 = atoi(NEXT);
		else if(a == "--dump-reads") dumpf = NEXT;      // reader only: name, bases, seed, filter verdict per read (tests diff it with the product's reader)
		else { fprintf(stderr, "cf_oracle: unknown option %s\n", a.c_str()); return 1; }
		#undef NEXT
	}
	if(p.min_hitlen < 15) p.min_hitlen = 15;               // centrifuge.cpp:1401-1407
	p.host_taxids = host.data(); p.n_host = host.size(); p.excluded_taxids = excl.data(); p.n_excluded = excl.size();
	char err[256];
	cfo_index* ix = cfo_index_load(idx.c_str(), err, sizeof err);
	if(!ix) { fprintf(stderr, "cf_oracle: %s\n", err); return 1; }
	bool paired = !m1.empty();
	FileBuf fa((paired ? m1 : u).c_str()); FileBuf* fbp = paired ? new FileBuf(m2.c_str()) : NULL;
	if(!fa.f || (paired && !fbp->f)) { fprintf(stderr, "cf_oracle: cannot open reads\n"); return 1; }
	FILE* fo = out == "-" ? stdout : fopen(out.c_str(), "wb");
	if(!fo) { fprintf(stderr, "cf_oracle: cannot open output\n"); return 1; }
	fputs("readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches\n", fo);
	cfo_stats st; memset(&st, 0, sizeof st);
	Ctx cx(*ix, p, &st);
	Species sp;
	bool firstA = true, firstB = true; u64 cntA = 0, cntB = 0;
	RawRead ra, rb; std::vector<cfo_rec> recs; std::string line;
	while(true) {
		bool okA, emptyA = false, emptyB = false;
		okA = fasta ? read_fasta(fa, ra, cntA, firstA, emptyA, trim5, trim3) : read_fastq(fa, ra, cntA, firstA, trim5, trim3);
		if(paired) {      // both mates are always attempted; a file running out first is an error (pat.cpp:330-420)
			const bool okB = fasta ? read_fasta(*fbp, rb, cntB, firstB, emptyB, trim5, trim3) : read_fastq(*fbp, rb, cntB, firstB, trim5, trim3);
			if(!okA && okB) { fprintf(stderr, "Error, fewer reads in file specified with -1 than in file specified with -2\n"); exit(1); }
			if(okA && !okB) { fprintf(stderr, "Error, fewer reads in file specified with -2 than in file specified with -1\n"); exit(1); }
			if(okB) cntB++;
		}
		if(!okA) break;
		cntA++;
		u32 seedA = gen_rand_seed(ra, 0), seedB = paired ? gen_rand_seed(rb, 0) : 0;
		bool pair = paired && !rb.seq.empty();
		bool f1 = n_filter(ra.seq) && ra.seq.size() >= 2, f2 = pair ? (n_filter(rb.seq) && rb.seq.size() >= 2) : false;
		if(!dumpf.empty()) {
			static FILE* df = NULL; if(!df) df = fopen(dumpf.c_str(), "wb");
			if(df) { fputs(ra.name.c_str(), df); fputc('\t', df); for(size_t q = 0; q < ra.seq.size(); q++) fputc("ACGTN"[ra.seq[q]], df); fprintf(df, "\t%u\t%d\n", seedA, f1 ? 1 : 0); fflush(df); }
			continue;
		}
		Rng rnd; rnd.init((f1 && f2) ? (seedA ^ seedB) : seedA);   // centrifuge.cpp:2609-2613
		recs.clear();
		int64_t max_score = 0;
		if(f1 || f2) {
			Unit un; setup_unit(un, ra.seq.data(), (u32)ra.seq.size(), pair ? rb.seq.data() : NULL, pair ? (u32)rb.seq.size() : 0, (uint8_t)((f1 ? 1 : 0) | (f2 ? 2 : 0)));
			classify_unit(cx, un, recs); st.reads++;
			for(int k = 0; k < un.nm; k++) { u64 L = un.m[k].len; max_score += (L > 15 ? (int64_t)((L - 15) * (L - 15)) : 0); }   // classifier.h:530-535
		}
		// AlnRes list; empty => the "unclassified" record (classifier.h:619-626), max_score 0
		struct R { int64_t score, max_score; u64 taxid; u32 hitlen, uid; bool uncl; };
		std::vector<R> rs;
		for(size_t k = 0; k < recs.size(); k++) { R r = {(int64_t)recs[k].score, max_score, recs[k].taxid, recs[k].hitlen, recs[k].uid, false}; rs.push_back(r); }
		if(rs.empty()) { R r = {0, 0, 0, 0, 0xFFFFFFFFu, true}; rs.push_back(r); }
		// AlnSetSumm::init aligner_result.h:398-427
		const int64_t INVALID = std::numeric_limits<int64_t>::min();
		int64_t best = INVALID, secbest = INVALID;
		for(size_t k = 0; k < rs.size(); k++) { int64_t sc = rs[k].score; if(sc > best) { secbest = best; best = sc; } else if(sc > secbest) secbest = sc; }
		// selectByScore aln_sink.h:1861-1927
		size_t sz = rs.size(); u64 num = std::min<u64>(sz, (u64)p.khits);
		std::vector<std::pair<int64_t, size_t> > buf(sz);
		for(size_t k = 0; k < sz; k++) buf[k] = std::make_pair(rs[k].score, k);
		std::sort(buf.begin(), buf.end()); std::reverse(buf.begin(), buf.end());
		size_t streak = 0;
		#define SHUF(begin, n) do { size_t left = (n); for(size_t q = (begin); q < (begin) + (n) - 1; q++) { u32 ri = rnd.next() % left; if(ri > 0) std::swap(buf[q], buf[q + ri]); left--; } } while(0)
		for(size_t k = 1; k < buf.size(); k++) {
			if(buf[k].first == buf[k - 1].first) { if(streak == 0) streak = 1; streak++; }
			else { if(streak > 1) SHUF(k - streak, streak); streak = 0; }
		}
		if(streak > 1) SHUF(buf.size() - streak, streak);
		#undef SHUF
		std::vector<size_t> select(num);
		for(size_t k = 0; k < num; k++) select[k] = buf[k].second;
		for(size_t k = 0; k + 1 < select.size(); k++) if(buf[k].first != buf[k + 1].first) { select.resize(k + 1); break; }
		u64 qlen = ra.seq.size() + (pair ? rb.seq.size() : 0);
		for(size_t k = 0; k < select.size(); k++) {
			const R& r = rs[select[k]];
			line.clear();
			append_read_id(line, ra.name); line.push_back('\t');
			// appendSeqID aln_sink.h:2220-2234 + uid choice classifier.h:557
			bool leaf = true; uint8_t taxRank = RANK_UNKNOWN;
			std::map<u64, TaxNode>::const_iterator t = ix->tree.find(r.taxid);
			if(t != ix->tree.end()) { leaf = t->second.leaf; taxRank = t->second.rank; }
			if(r.uncl) line += "unclassified";
			else if(leaf) line += (r.uid != 0xFFFFFFFFu ? ix->uid_to_tid[r.uid].first.c_str() : rank_string(taxRank));
			else line += rank_string(taxRank);
			line.push_back('\t'); append_taxid(line, r.taxid);
			char b[128];
			snprintf(b, sizeof b, "\t%llu\t%llu\t%llu\t%llu\t%llu\n", (unsigned long long)r.score,
			         (unsigned long long)(secbest != INVALID ? secbest : 0), (unsigned long long)r.hitlen,
			         (unsigned long long)qlen, (unsigned long long)select.size());
			line += b;
			fputs(line.c_str(), fo);
			sp.add(r.taxid, r.score, r.max_score, (u32)select.size());
		}
	}
	if(fo != stdout) fclose(fo);
	delete fbp;
	if(!report.empty()) {                                    // centrifuge.cpp:3231-3319
		if(abundance) calc_abundance(*ix, sp);
		std::ofstream ro(report.c_str());
		ro << "name\ttaxID\ttaxRank\tgenomeSize\tnumReads\tnumUniqueReads\tabundance" << std::endl;
		for(std::map<u64, ReadCounts>::const_iterator it = sp.counts.begin(); it != sp.counts.end(); ++it) {
			u64 taxid = it->first; if(taxid == 0) continue;
			std::map<u64, std::string>::const_iterator nm = ix->name.find(taxid);
			if(nm != ix->name.end()) ro << nm->second; else ro << taxid;
			ro << '\t' << taxid << '\t';
			uint8_t rank = 0; bool leaf = false;
			std::map<u64, TaxNode>::const_iterator t = ix->tree.find(taxid);
			if(t != ix->tree.end()) { rank = t->second.rank; leaf = t->second.leaf; }
			if(rank == RANK_UNKNOWN && leaf) ro << "leaf"; else ro << rank_string(rank);
			ro << '\t';
			std::map<u64, u64>::const_iterator s = ix->size.find(taxid);
			ro << (s != ix->size.end() ? s->second : 0) << '\t' << it->second.n_reads << '\t' << it->second.n_unique_reads << '\t';
			std::map<u64, double>::const_iterator ab = sp.abundance_len.find(taxid);
			if(ab != sp.abundance_len.end()) ro << ab->second; else ro << "0.0";
			ro << std::endl;
		}
	}
	if(!statsf.empty()) {
		FILE* sf = fopen(statsf.c_str(), "w");
		if(sf) {
			fprintf(sf, "{\"reads\": %llu, \"partial_searches\": %llu, \"ftab_probes\": %llu, \"lf_range_steps\": %llu, "
			        "\"lf_range_same_side\": %llu, \"lf_single_steps\": %llu, \"sides_search\": %llu, \"walk_steps\": %llu, "
			        "\"rows_resolved\": %llu, \"hits_resolved\": %llu, \"ext_searches\": %llu, \"sample_width\": %d}\n",
			        (unsigned long long)st.reads, (unsigned long long)st.partial_searches, (unsigned long long)st.ftab_probes,
			        (unsigned long long)st.lf_range_steps, (unsigned long long)st.lf_range_same_side, (unsigned long long)st.lf_single_steps,
			        (unsigned long long)st.sides_search, (unsigned long long)st.walk_steps, (unsigned long long)st.rows_resolved,
			        (unsigned long long)st.hits_resolved, (unsigned long long)st.ext_searches, ix->offw ? 4 : 2);
			fclose(sf);
		}
	}
	cfo_index_free(ix);
	return 0;
}

#ifdef CFO_MAIN
int main(int argc, const char** argv) { return cfo_main(argc, argv); }
#endif
