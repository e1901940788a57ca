// cf_logic.h -- per-read classification logic shared by the CUDA kernels (device) and the
// CPU unit tests (host compilation, tests/ only; the product never runs this on the host).
//
// Everything here is scalar, branchy, integer code: the parts of Classifier::go that are not
// the bandwidth-bound FM walks.  The warp-cooperative walks live in cf_kernels.cu.
// Reference behaviour restated (paths relative to the reference tree):
//   partialSearch            hi_aligner.h:903-1031
//   searchForwardAndReverse  classifier.h:646-896 (extend / twin removal / trim parts)
//   getForwardOrReverseHit   classifier.h:898-941
//   compareBWTHits + sort    classifier.h:267,1058-1086, ds.h:775-778 (std::sort, libstdc++)
//   go(): resolve loop, hit map, finalize, host rule, tree reduction, emit  classifier.h:212-571
#ifndef CF_LOGIC_H_
#define CF_LOGIC_H_

#include <stdint.h>

#ifdef __CUDACC__
#define CFB_HD __host__ __device__ __forceinline__
#define CFB_HDN __host__ __device__
#else
#define CFB_HD inline
#define CFB_HDN inline
#endif

namespace cfb {

static const uint64_t kOff = 0xffffffffffffffffull;
static const uint64_t kUnk = 0xfffffffffffffffeull;   // HitRec.top == bot == kUnk: a hit shorter than min_hitlen whose SA range was not computed (death-depth table)
static const uint32_t kBwNone = 0xffffffffu;       // BWTHit::reset(): _bwoff = OFF_MASK

// ----------------------------------------------------------------------------------------
// Device/host view of the index.  Only lineRate 7 (128-byte sides: 96 B of 2-bit BWT = 384
// bases, then occ[A,C,G,T] as 4 x u64 counted before the side) is supported by the kernels.
// ----------------------------------------------------------------------------------------
struct IndexView {
	const uint64_t* sides;      // num_sides * 16 u64 (128-byte aligned on device)
	const uint64_t* ftab;
	const uint64_t* eftab;
	const uint16_t* sample16;   // exactly one of sample16/sample32 is non-null
	const uint32_t* sample32;
	const uint64_t* brow;       // sorted boundary rows
	const uint32_t* bseq;
	const uint32_t* bbits;      // prefilter bitmap over row >> bshift
	const uint64_t* seq_taxid;
	const int32_t*  seq_path;
	const uint64_t* paths;      // n_paths * 10
	const uint8_t*  seq_excluded;   // per-sequence flag (per ctx; may be null)
	const uint64_t* host_taxids;    // sorted expanded host set (per ctx)
	const uint64_t* blocks;         // device-only re-blocked FM index: 64-byte blocks of 128 rows (cfb200.cu)
	const uint64_t* rankv;          // device-only per-base rank sectors: 32 bytes per (192 rows, base)
	const uint64_t* rank16;         // device-only rank entries: 16 bytes (occ_c, 64 indicator bits) per (64 rows, base)
	const uint64_t* ftab2;          // device-only fused ftab: (top, bot) per 10-mer, eftab already resolved
	const uint16_t* rtab16;         // device-only resolve table: sequence id of EVERY SA row (one of rtab16/rtab32, or neither)
	const uint32_t* rtab32;
	const uint64_t* ftabk;          // device-only extended jump table: (top, bot) per K-mer, K = ftabk_chars (0 = absent)
	const uint64_t* walk8;          // device-only: per SA row, the row 8 LF steps on | the 8 BWT bases met << 40 | #valid steps << 56 (null = absent)
	uint64_t walk8_rows;            // rows [0, walk8_rows) have a walk8 entry (the table may cover a prefix of the rows when HBM is short)
	const uint8_t*  ftabd;          // device-only "death depth" table: 2 bits per (D = base K + 3)-mer, see k_build_ftabd (null = absent)
	int32_t  ftabd_chars, ftabd_base;   // D and the K of the jump table it extends (ftabk_chars, or ftab_chars without a K-mer table)
	uint64_t len, zoff, zside, fchr[4], last_boundary, num_sides, num_blocks;
	uint32_t zoffc, n_boundaries, n_seqs, n_host;
	int32_t  off_rate, ftab_chars, bshift, ftabk_chars;
};

struct Params {
	uint32_t khits, min_hitlen, ihits, increment;
	uint32_t tree_traverse, class_rank_slot;
};

struct HitRec { uint64_t top, bot; uint32_t bwoff, len; };   // 24 bytes

// A row queued for resolution carries the plan of the scoring pass in its high bits, so that pass reads one
// contiguous array instead of gathering the hit lists again: the first row of every counted hit has kRowStart set
// plus the hit length (bits 40..55), the list it came from (mate bit 56, strand bit 57) and kRowSameTs when the hit
// carries the same time stamp as the previous counted hit.  That happens in the reference when a list loop is left
// through its `break` (the loop-header ts++ is skipped, classifier.h:283-345) and the next list starts with a counted
// hit: the second hit then does not add its score to entries the first one touched -- a quirk the output depends on.
// SA rows need 40 bits.
static const uint64_t kRowMask = (1ull << 40) - 1ull, kRowStart = 1ull << 63, kRowSameTs = 1ull << 58;

struct Counters {   // algorithmic-operation counters (SURVEY.md section 8d definition)
	unsigned long long units, partial_searches, ftab_probes, sides_search, walk_steps, rows_resolved, lf_steps, ext_searches;
	// the product's own load requests (count mode 2: every derived table live, one entry per lane-level gather):
	// rank16 entries (16 B), 10-mer table entries (16 B), K-mer table entries (16 B), walk8 entries (8 B)
	unsigned long long req_rank16, req_ftab2, req_ftabk, req_walk8, req_ftabd;
};

CFB_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}

// positions (bit 2i) where the 2-bit code at i equals c
CFB_HD uint64_t match2(uint64_t w, int c) {
	uint64_t y = ~(w ^ ((uint64_t)c * 0x5555555555555555ull));
	return y & (y >> 1) & 0x5555555555555555ull;
}

CFB_HD int bwt_char(const IndexView& v, uint64_t row) {
#ifdef __CUDA_ARCH__
	if(v.rank16) {     // device replica: the indicator bits of the row's 64-row block (one 64-byte chunk); the '$' row has no bit and reads as A, as the file stores it
		const uint64_t* e = v.rank16 + (row >> 6) * 8;
		const uint32_t o = (uint32_t)(row & 63);
		return (int)(((e[3] >> o) & 1ull) * 1 + ((e[5] >> o) & 1ull) * 2 + ((e[7] >> o) & 1ull) * 3);
	}
#endif
	uint64_t s = row / 384; uint32_t off = (uint32_t)(row - s * 384);
	return (int)((v.sides[s * 16 + (off >> 5)] >> ((off & 31) * 2)) & 3);
}

// LF(row, c) = fchr[c] + occ_side[c] + rank_c(side, off) - ['$' stored as A lies before off]
// (countBt2Side bt2_idx.h:2192-2227, countUpTo :2364-2425).  Scalar version: one thread reads
// the words it needs.
CFB_HD uint64_t lf_scalar(const IndexView& v, uint64_t row, int c) {
#ifdef __CUDA_ARCH__
	if(v.rank16) {     // device replica: one 16-byte rank16 entry (occ before the block, '$' excluded | indicator bits), same value as below
		const uint64_t* e = v.rank16 + ((row >> 6) * 4 + (uint64_t)c) * 2;
		return v.fchr[c] + (e[0] & 0x7fffffffffffffffull) + (uint64_t)popc64(e[1] & (((uint64_t)1 << (row & 63)) - 1));
	}
#endif
	uint64_t s = row / 384; uint32_t off = (uint32_t)(row - s * 384);
	const uint64_t* w = v.sides + s * 16;
	uint32_t full = off >> 5, rem = off & 31;
	uint64_t n = 0;
	for(uint32_t k = 0; k < full; k++) n += (uint64_t)popc64(match2(w[k], c));
	if(rem) n += (uint64_t)popc64(match2(w[full], c) & (((uint64_t)1 << (2 * rem)) - 1));
	if(c == 0 && s == v.zside && v.zoffc < off) n--;
	return v.fchr[c] + w[12 + c] + n;
}

// strand sequence access without materialising the reverse complement:
// strand 0: seq[j] = fw[j]; strand 1: seq[j] = comp(fw[len-1-j])  (Read::constructRevComps)
CFB_HD int seq_at(const uint8_t* fw, uint32_t len, int strand, uint32_t j) {
	if(strand == 0) return fw[j];
	int c = fw[len - 1 - j];
	return c > 3 ? 4 : 3 - c;
}

CFB_HD uint64_t ftab_hi(const IndexView& v, uint64_t e) { return e <= v.len ? e : v.eftab[(e ^ kOff) * 2 + 1]; }
CFB_HD uint64_t ftab_lo(const IndexView& v, uint64_t e) { return e <= v.len ? e : v.eftab[(e ^ kOff) * 2]; }

// One partialSearch from offset `cur` (scalar; used by the rare extend step and by CPU tests).
// Returns the hit it would append and the new cur / done flags.
CFB_HDN void partial_search_scalar(const IndexView& v, const uint8_t* fw, uint32_t len, int strand,
                                   uint32_t cur, HitRec& out, uint32_t& new_cur, bool& done, Counters* ctr) {
	const uint32_t fc = (uint32_t)v.ftab_chars;
	uint32_t offset = cur, dep = cur;
	done = false;
	if(ctr) ctr->partial_searches++;
	if(len - dep < fc) {
		new_cur = len; out.top = out.bot = kOff; out.bwoff = offset; out.len = len - offset; done = true; return;
	}
	for(uint32_t i = 0; i < fc; i++) {
		if(seq_at(fw, len, strand, len - dep - 1 - i) > 3) {
			new_cur = cur + i + 1; out.top = out.bot = kOff; out.bwoff = offset; out.len = new_cur - offset;
			done = new_cur >= len; return;
		}
	}
	uint64_t fi = 0;
	for(uint32_t i = 0; i < fc; i++) fi = (fi << 2) | (uint64_t)seq_at(fw, len, strand, len - dep - fc + i);
	uint64_t top = ftab_hi(v, v.ftab[fi]), bot = ftab_lo(v, v.ftab[fi + 1]);
	if(ctr) ctr->ftab_probes++;
	dep += fc;
	if(bot <= top) {
		new_cur = dep; out.top = out.bot = kOff; out.bwoff = offset; out.len = dep - offset; done = dep >= len; return;
	}
	while(dep < len) {
		int c = seq_at(fw, len, strand, len - dep - 1);
		if(c > 3) break;
		uint64_t t, b;
		if(bot - top != 1) {
			t = lf_scalar(v, top, c); b = lf_scalar(v, bot, c);
			if(ctr) { ctr->lf_steps += 2; ctr->sides_search += ((top % 384) + (bot - top) < 384) ? 1 : 2; }
		} else {
			if(ctr) { ctr->lf_steps += 1; ctr->sides_search += 1; }
			if(bwt_char(v, top) != c || top == v.zoff) break;
			t = lf_scalar(v, top, c); b = t + 1;
		}
		if(b <= t) break;
		top = t; bot = b; dep++;
	}
	out.top = top; out.bot = bot; out.bwoff = offset; out.len = dep - offset;
	new_cur = dep; done = dep >= len;
}

// Whole greedy search of one strand (what the search kernel computes cooperatively).
// Used on the host by tests to cross-check kernel output stage by stage.
CFB_HDN uint32_t search_strand_scalar(const IndexView& v, const Params& p, const uint8_t* fw, uint32_t len, int strand,
                                      HitRec* hits, uint32_t cap, Counters* ctr) {
	uint32_t cur = 0, n = 0;
	if(len == 0) return 0;
	for(;;) {
		HitRec h; uint32_t nc; bool done;
		partial_search_scalar(v, fw, len, strand, cur, h, nc, done, ctr);
		if(n < cap) hits[n] = h;
		n++;
		cur = nc;
		if(done) break;
		if(h.len > p.increment) cur += 1;
		if(cur + p.min_hitlen >= len) break;
	}
	return n;
}

CFB_HD uint64_t bw64(const HitRec& h) { return h.bwoff == kBwNone ? kOff : (uint64_t)h.bwoff; }
CFB_HD uint64_t hsize(const HitRec& h) { return h.bot - h.top; }

// Post-search part of searchForwardAndReverse for one mate: extend, twin removal, trim.
CFB_HDN void post_search(const IndexView& v, const Params& p, const uint8_t* fw, uint32_t rdlen,
                         HitRec* F, uint32_t nF, HitRec* R, uint32_t nR, Counters* ctr) {
	const uint64_t minHitLen = p.min_hitlen;
	uint64_t sum[2] = {0, 0};
	for(uint32_t i = 0; i < nF; i++) if(F[i].len >= minHitLen) sum[0] += F[i].len;
	for(uint32_t i = 0; i < nR; i++) if(R[i].len >= minHitLen) sum[1] += R[i].len;
	if(sum[0] >= minHitLen && sum[1] >= minHitLen) {
		for(uint32_t i = 0; i < nF; i++) {
			const uint64_t len = F[i].len, l = bw64(F[i]), r = l + len;     // locals are not refreshed (classifier.h:795-798)
			for(uint32_t j = 0; j < nR; j++) {
				const uint64_t rclen = R[j].len;
				if(len < minHitLen && rclen < minHitLen) continue;
				const uint64_t rc_l = (uint64_t)rdlen - bw64(R[j]) - R[j].len, rc_r = rc_l + rclen;
				if(r <= rc_l) continue;
				if(rc_r <= l) continue;
				if(l == rc_l && r == rc_r) continue;
				if(l < rc_l && r > rc_r) continue;
				if(l > rc_l && r < rc_r) continue;
				if(l > rc_l) {
					HitRec t; uint32_t nc; bool dn;
					if(ctr) ctr->ext_searches++;
					partial_search_scalar(v, fw, rdlen, 0, (uint32_t)rc_l, t, nc, dn, ctr);
					if((uint64_t)t.len == len + l - rc_l) F[i] = t;
				}
				if(r > rc_r) {
					HitRec t; uint32_t nc; bool dn;
					if(ctr) ctr->ext_searches++;
					partial_search_scalar(v, fw, rdlen, 1, (uint32_t)((uint64_t)rdlen - r), t, nc, dn, ctr);
					if((uint64_t)t.len == rclen + r - rc_r) R[j] = t;
				}
			}
		}
		for(uint32_t i = 0; i < nF; i++) {
			const uint64_t len = F[i].len, l = bw64(F[i]), r = l + len;
			for(uint32_t j = 0; j < nR; j++) {
				const uint64_t rclen = R[j].len;
				const uint64_t rc_l = (uint64_t)rdlen - bw64(R[j]) - R[j].len, rc_r = rc_l + rclen;
				if(rc_l < l) break;
				if(len != rclen) continue;
				if(l == rc_l && r == rc_r && hsize(F[i]) + hsize(R[j]) > (uint64_t)p.ihits) {
					F[i].top = F[i].bot = 0; F[i].bwoff = kBwNone; F[i].len = 0;
					R[j].top = R[j].bot = 0; R[j].bwoff = kBwNone; R[j].len = 0;
					break;
				}
			}
		}
	}
	for(int s = 0; s < 2; s++) {
		HitRec* L = s == 0 ? F : R; const uint32_t n = s == 0 ? nF : nR;
		if(n < 2) continue;
		for(uint32_t i = 0; i + 1 < n; i++) {
			for(uint32_t j = i + 1; j < n; j++) {
				const uint64_t abw = bw64(L[i]), bbw = bw64(L[j]);
				if(abw >= bbw) { L[i].len = 0; break; }
				if(abw + L[i].len <= bbw) break;
				if(L[i].len >= L[j].len) {
					const uint64_t e = bbw + L[j].len, nb = abw + L[i].len;
					L[j].bwoff = (uint32_t)nb; L[j].len = (uint32_t)(e - nb);
				} else L[i].len = (uint32_t)(bbw - abw);
			}
		}
	}
}

// strand choice: returns lo | hi<<1 style pair as (first, second)
CFB_HD void choose_strand(const Params& p, const HitRec* F, uint32_t nF, const HitRec* R, uint32_t nR, int& first, int& second) {
	uint64_t avg[2] = {0, 0}, mx[2] = {0, 0};
	for(int s = 0; s < 2; s++) {
		const HitRec* L = s == 0 ? F : R; const uint32_t n = s == 0 ? nF : nR;
		for(uint32_t i = 0; i < n; i++) {
			const uint64_t len = L[i].len;
			if(len < p.min_hitlen) continue;
			avg[s] += (len - 15) * (len - 15);
			if(len > mx[s]) mx[s] = len;
		}
	}
	if(avg[0] != avg[1]) { first = avg[0] > avg[1] ? 0 : 1; second = first + 1; return; }
	if(mx[0] != mx[1])   { first = mx[0] > mx[1] ? 0 : 1;   second = first + 1; return; }
	first = 0; second = 2;
}

struct HitLess {   // compareBWTHits classifier.h:1058-1086
	CFB_HD bool operator()(const HitRec& a, const HitRec& b) const {
		const uint64_t al = a.len, bl = b.len, as = hsize(a), bs = hsize(b);
		if(al >= 22 || bl >= 22) {
			if(al >= 22 && bl >= 22) { if(as < bs) return true; if(as > bs) return false; }
			if(bl < al) return true;
			if(bl > al) return false;
		}
		if(bl * as < al * bs) return true;
		if(bl * as > al * bs) return false;
		if(as < bs) return true;
		if(as > bs) return false;
		if(bl < al) return true;
		if(bl > al) return false;
		return false;
	}
};

// ----------------------------------------------------------------------------------------
// std::sort as implemented by libstdc++ (bits/stl_algo.h, bits/stl_heap.h): introsort with
// median-of-3 pivot, threshold 16, depth limit 2*floor(log2 n), heapsort fallback, final
// insertion sort.  Restated because the reference's result depends on the exact permutation
// this algorithm produces when the comparator reports ties.
// ----------------------------------------------------------------------------------------
template <class T, class C> CFB_HD void ss_swap(T& a, T& b, C&) { T t = a; a = b; b = t; }

template <class T, class C> CFB_HDN void ss_unguarded_linear_insert(T* last, C comp) {
	T val = *last; T* next = last - 1;
	while(comp(val, *next)) { *last = *next; last = next; --next; }
	*last = val;
}
template <class T, class C> CFB_HDN void ss_insertion_sort(T* first, T* last, C comp) {
	if(first == last) return;
	for(T* i = first + 1; i != last; ++i) {
		if(comp(*i, *first)) { T val = *i; for(T* q = i; q != first; --q) *q = *(q - 1); *first = val; }
		else ss_unguarded_linear_insert(i, comp);
	}
}
template <class T, class C> CFB_HDN void ss_push_heap(T* first, long hole, long top, T value, C comp) {
	long parent = (hole - 1) / 2;
	while(hole > top && comp(first[parent], value)) { first[hole] = first[parent]; hole = parent; parent = (hole - 1) / 2; }
	first[hole] = value;
}
template <class T, class C> CFB_HDN void ss_adjust_heap(T* first, long hole, long len, T value, C comp) {
	const long top = hole; long child = hole;
	while(child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if(comp(first[child], first[child - 1])) child--;
		first[hole] = first[child]; hole = child;
	}
	if((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); first[hole] = first[child - 1]; hole = child - 1; }
	ss_push_heap(first, hole, top, value, comp);
}
template <class T, class C> CFB_HDN void ss_heapsort(T* first, T* last, C comp) {   // __partial_sort(first,last,last)
	long len = (long)(last - first);
	if(len >= 2) {
		long parent = (len - 2) / 2;
		for(;;) { T v = first[parent]; ss_adjust_heap(first, parent, len, v, comp); if(parent == 0) break; parent--; }
	}
	while(last - first > 1) { --last; T v = *last; *last = *first; ss_adjust_heap(first, 0L, (long)(last - first), v, comp); }
}
template <class T, class C> CFB_HDN T* ss_partition_pivot(T* first, T* last, C comp) {
	T* mid = first + (last - first) / 2;
	T *a = first + 1, *b = mid, *c = last - 1;
	if(comp(*a, *b)) { if(comp(*b, *c)) ss_swap(*first, *b, comp); else if(comp(*a, *c)) ss_swap(*first, *c, comp); else ss_swap(*first, *a, comp); }
	else if(comp(*a, *c)) ss_swap(*first, *a, comp);
	else if(comp(*b, *c)) ss_swap(*first, *c, comp);
	else ss_swap(*first, *b, comp);
	T* lo = first + 1; T* hi = last; T* pivot = first;
	for(;;) {
		while(comp(*lo, *pivot)) ++lo;
		--hi;
		while(comp(*pivot, *hi)) --hi;
		if(!(lo < hi)) return lo;
		ss_swap(*lo, *hi, comp);
		++lo;
	}
}
template <class T, class C> CFB_HDN void std_sort(T* first, T* last, C comp) {
	const long n = (long)(last - first);
	if(n <= 0) return;
	if(n > 16) {
		int lg = 0; for(long t = n; t > 1; t >>= 1) lg++;
		// explicit stack instead of recursion: sub-ranges are disjoint, so order is irrelevant
		T* st_first[64]; T* st_last[64]; int st_depth[64]; int sp = 0;
		st_first[0] = first; st_last[0] = last; st_depth[0] = 2 * lg; sp = 1;
		while(sp > 0) {
			--sp; T* f = st_first[sp]; T* l = st_last[sp]; int depth = st_depth[sp];
			while(l - f > 16) {
				if(depth == 0) { ss_heapsort(f, l, comp); break; }
				--depth;
				T* cut = ss_partition_pivot(f, l, comp);
				st_first[sp] = cut; st_last[sp] = l; st_depth[sp] = depth; sp++;
				l = cut;
			}
		}
		ss_insertion_sort(first, first + 16, comp);
		for(T* i = first + 16; i != last; ++i) ss_unguarded_linear_insert(i, comp);
	} else ss_insertion_sort(first, last, comp);
}

// ----------------------------------------------------------------------------------------
// Resolve planning + scoring
// ----------------------------------------------------------------------------------------
struct UnitHits {      // hit lists of one unit: [mate][strand]
	HitRec* L[2][2]; uint32_t n[2][2]; uint32_t rdlen[2]; int n_mates;
};

// Visit order and per-visit maxGenomeHitSize of go() (classifier.h:228,243-265); calls
// f(rdi, fwi, maxG) for each visited strand list, in order.  Lists must already be post_search'ed.
template <class F> CFB_HDN void for_each_visit(const Params& p, const UnitHits& u, F& f) {
	uint64_t maxG = p.khits;
	for(int rdi = 0; rdi < u.n_mates; rdi++) {
		int a, b;
		choose_strand(p, u.L[rdi][0], u.n[rdi][0], u.L[rdi][1], u.n[rdi][1], a, b);
		for(int fwi = a; fwi < b; fwi++) {
			const HitRec* L = u.L[rdi][fwi]; const uint32_t n = u.n[rdi][fwi];
			for(uint32_t hi = 0; hi < n; hi++) if(L[hi].len >= p.min_hitlen && hsize(L[hi]) > maxG) maxG = hsize(L[hi]);
			if(maxG > p.khits) maxG += p.khits;
			f(rdi, fwi, maxG);
		}
	}
}

struct SortAndCount {   // prep pass: sort each visited list, count rows to resolve
	const Params& p; UnitHits& u; uint64_t rows;
	CFB_HD SortAndCount(const Params& p_, UnitHits& u_) : p(p_), u(u_), rows(0) {}
	CFB_HD void operator()(int rdi, int fwi, uint64_t maxG) {
		HitRec* L = u.L[rdi][fwi]; const uint32_t n = u.n[rdi][fwi];
		std_sort(L, L + n, HitLess());
		uint64_t cnt = 0;
		for(uint32_t hi = 0; hi < n; hi++) {
			if(L[hi].len <= p.min_hitlen) continue;
			if(hsize(L[hi]) == 0) continue;
			const uint64_t nelt = hsize(L[hi]) < maxG ? hsize(L[hi]) : maxG;
			if(nelt > p.ihits) continue;           // resolved-then-discarded in the reference (classifier.h:299)
			rows += nelt; cnt += nelt;
			if(cnt >= maxG) break;
		}
	}
};

struct CountRows {      // same count as SortAndCount on lists that are already sorted (re-run after a capacity overflow)
	const Params& p; const UnitHits& u; uint64_t rows;
	CFB_HD CountRows(const Params& p_, const UnitHits& u_) : p(p_), u(u_), rows(0) {}
	CFB_HD void operator()(int rdi, int fwi, uint64_t maxG) {
		const HitRec* L = u.L[rdi][fwi]; const uint32_t n = u.n[rdi][fwi];
		uint64_t cnt = 0;
		for(uint32_t hi = 0; hi < n; hi++) {
			if(L[hi].len <= p.min_hitlen) continue;
			if(hsize(L[hi]) == 0) continue;
			const uint64_t nelt = hsize(L[hi]) < maxG ? hsize(L[hi]) : maxG;
			if(nelt > p.ihits) continue;
			rows += nelt; cnt += nelt;
			if(cnt >= maxG) break;
		}
	}
};

struct EmitRows {       // second pass: write the SA rows to resolve, in consumption order, with the scoring plan
	const Params& p; const UnitHits& u; uint64_t* out; uint64_t k; uint32_t ts, last_ts;
	CFB_HD EmitRows(const Params& p_, const UnitHits& u_, uint64_t* o) : p(p_), u(u_), out(o), k(0), ts(0), last_ts(0xffffffffu) {}
	CFB_HD void operator()(int rdi, int fwi, uint64_t maxG) {
		const HitRec* L = u.L[rdi][fwi]; const uint32_t n = u.n[rdi][fwi];
		uint64_t cnt = 0;
		for(uint32_t hi = 0; hi < n; hi++, ts++) {            // ts advances exactly as in the reference's loop header
			if(L[hi].len <= p.min_hitlen) continue;
			if(hsize(L[hi]) == 0) continue;
			const uint64_t nelt = hsize(L[hi]) < maxG ? hsize(L[hi]) : maxG;
			if(nelt > p.ihits) continue;
			const uint64_t head = kRowStart | ((uint64_t)(L[hi].len & 0xffffu) << 40) | ((uint64_t)(rdi & 1) << 56) | ((uint64_t)(fwi & 1) << 57)
			                    | (ts == last_ts ? kRowSameTs : 0ull);
			last_ts = ts;
			for(uint64_t e = 0; e < nelt; e++) out[k++] = ((L[hi].top + e) & kRowMask) | (e == 0 ? head : 0ull);
			cnt += nelt;
			if(cnt >= maxG) break;
		}
	}
};

struct Entry {          // HitCount classifier.h:31-57 (fields that influence output)
	uint64_t uniqueID, taxID;
	uint32_t scores[2][2], lens[2][2];
	uint32_t score, hitlen, ts;
	int32_t  pid;        // path id or -1 (empty path)
	uint8_t  rank;
	uint8_t  pad[3];
};
struct TaxCnt { uint32_t count; uint32_t pad; uint64_t parent; };
struct TaxCntLess { CFB_HD bool operator()(const TaxCnt& a, const TaxCnt& b) const { return a.count < b.count || (a.count == b.count && a.parent < b.parent); } };

struct OutRec { uint64_t taxid; uint32_t score, hitlen, uid, pad; };   // == cfb_rec

CFB_HD bool is_host(const IndexView& v, uint64_t taxid) {
	uint32_t lo = 0, hi = v.n_host;
	while(lo < hi) { uint32_t mid = (lo + hi) >> 1; if(v.host_taxids[mid] < taxid) lo = mid + 1; else hi = mid; }
	return lo < v.n_host && v.host_taxids[lo] == taxid;
}
CFB_HD uint32_t path_size(const Entry& e) { return e.pid < 0 ? 0u : 10u; }
CFB_HD uint64_t path_at(const IndexView& v, const Entry& e, uint32_t i) { return v.paths[(uint64_t)e.pid * 10 + i]; }

// Third pass: consume the resolved ids along the plan carried by the row words, build the hit map
// (classifier.h:299-345 + addHitToHitMap :982-1050).  Time stamps are non-decreasing along the plan, so "same as the
// previous hit" (kRowSameTs) reproduces every equality the reference's counter produces.
CFB_HDN uint32_t score_plan(const IndexView& v, const Params& p, const uint64_t* rows, const uint32_t* ids, uint64_t n, Entry* map) {
	uint32_t nmap = 0, ts = 0;
	for(uint64_t k = 0; k < n;) {
		const uint64_t head = rows[k];
		if(!(head & kRowSameTs)) ts++;
		const uint64_t hl = (head >> 40) & 0xffffu; const int rdi = (int)((head >> 56) & 1), fwi = (int)((head >> 57) & 1);
		uint64_t nelt = 1;
		while(k + nelt < n && !(rows[k + nelt] & kRowStart)) nelt++;
		const uint32_t* my = ids + k; k += nelt;
		const uint32_t sc = (uint32_t)((hl - 15) * (hl - 15));
		for(uint64_t e = 0; e < nelt; e++) {
			const uint32_t ref = my[e];
			bool dup = false;                       // coord_ids de-duplication, first-seen order
			for(uint64_t q = 0; q < e; q++) if(my[q] == ref) { dup = true; break; }
			if(dup) continue;
			uint64_t taxID = ref < v.n_seqs ? v.seq_taxid[ref] : 0;
			if(v.seq_excluded && ref < v.n_seqs && v.seq_excluded[ref]) continue;
			const int32_t pid = ref < v.n_seqs ? v.seq_path[ref] : -1;
			uint8_t rank = (uint8_t)p.class_rank_slot;
			if(rank > 0 && pid >= 0) {
				for(; rank < 10; rank++) { const uint64_t t = v.paths[(uint64_t)pid * 10 + rank]; if(t != 0) { taxID = t; break; } }
			}
			uint32_t idx = 0;
			for(; idx < nmap; ++idx) {
				const bool same = rank == 0 ? ((uint64_t)ref == map[idx].uniqueID) : (taxID == map[idx].taxID);
				if(same) {
					if(map[idx].ts != ts) { map[idx].scores[rdi][fwi] += sc; map[idx].lens[rdi][fwi] += (uint32_t)hl; map[idx].ts = ts; }
					break;
				}
			}
			if(idx >= nmap) {
				Entry& n2 = map[nmap++];
				n2.uniqueID = ref; n2.taxID = taxID;
				n2.scores[0][0] = n2.scores[0][1] = n2.scores[1][0] = n2.scores[1][1] = 0;
				n2.lens[0][0] = n2.lens[0][1] = n2.lens[1][0] = n2.lens[1][1] = 0;
				n2.scores[rdi][fwi] = sc; n2.lens[rdi][fwi] = (uint32_t)hl;
				n2.score = 0; n2.hitlen = 0; n2.ts = ts; n2.pid = pid; n2.rank = rank;
			}
		}
	}
	return nmap;
}

// rank > 0 with an empty path: the reference's loop `for(; rank < path.size(); ...)` does not
// run and rank keeps its configured value; handled above because pid < 0 skips the loop.

// finalize + host rule + tree reduction + emit (classifier.h:380-571).
// map/nmap: hit map; tc: scratch of >= nmap TaxCnt; out: >= nmap records.  Returns #records.
CFB_HDN uint32_t reduce_and_emit(const IndexView& v, const Params& p, bool paired, Entry* map, uint32_t nmap,
                                 TaxCnt* tc, OutRec* out) {
	const uint32_t k = p.khits;
	for(uint32_t i = 0; i < nmap; i++) {
		Entry& h = map[i];
		const uint32_t s0 = h.scores[0][0] > h.scores[0][1] ? h.scores[0][0] : h.scores[0][1];
		const uint32_t l0 = h.lens[0][0] > h.lens[0][1] ? h.lens[0][0] : h.lens[0][1];
		if(paired) {
			const uint32_t s1 = h.scores[1][0] > h.scores[1][1] ? h.scores[1][0] : h.scores[1][1];
			const uint32_t l1 = h.lens[1][0] > h.lens[1][1] ? h.lens[1][0] : h.lens[1][1];
			h.score = s0 + s1; h.hitlen = l0 + l1;
		} else { h.score = s0; h.hitlen = l0; }
	}
	int64_t best = 0; bool only_host = false;
	for(uint32_t i = 0; i < nmap; i++) {
		if((int64_t)map[i].score > best) { best = map[i].score; only_host = is_host(v, map[i].taxID); }
		else if((int64_t)map[i].score == best) only_host |= is_host(v, map[i].taxID);
	}
	if(!only_host && nmap > k) {
		uint32_t bs = map[0].score;
		for(uint32_t i = 1; i < nmap; i++) if(bs < map[i].score) bs = map[i].score;
		for(int i = 0; i < (int)nmap; i++) {
			if(map[i].score < bs) { if(i + 1 < (int)nmap) map[i] = map[nmap - 1]; nmap--; i--; }
		}
		if(!p.tree_traverse && nmap > k) return 0;
		uint8_t rank = 0;
		while(nmap > k) {
			uint32_t ntc = 0;
			for(uint32_t i = 0; i < nmap; i++) {
				Entry& h = map[i];
				while(h.rank < rank) {
					if((uint32_t)h.rank + 1 >= path_size(h)) { h.rank = 255; break; }
					h.rank += 1; h.taxID = path_at(v, h, h.rank);
				}
				if(h.rank > rank) continue;
				const uint64_t parent = ((uint32_t)rank + 1 >= path_size(h)) ? 1 : path_at(v, h, rank + 1);
				if(parent == 0) continue;
				uint32_t j = 0;
				for(; j < ntc; j++) if(tc[j].parent == parent) { tc[j].count += 1; break; }
				if(j == ntc) { tc[ntc].count = 1; tc[ntc].pad = 0; tc[ntc].parent = parent; ntc++; }
			}
			if(ntc == 0) {
				if(rank < path_size(map[0])) { rank++; continue; } else break;
			}
			ss_heapsort(tc, tc + ntc, TaxCntLess());    // keys are distinct: any correct sort gives std::sort's result
			uint32_t j = ntc;
			while(j-- > 0) {
				const uint64_t parent = tc[j].parent;
				for(uint32_t i = 0; i < nmap; i++) {
					Entry& h = map[i];
					if(h.rank != rank) continue;
					const uint64_t cur_parent = ((uint32_t)rank + 1 >= path_size(h)) ? 1 : path_at(v, h, rank + 1);
					if(parent == cur_parent) { h.uniqueID = kOff; h.rank = rank + 1; h.taxID = parent; }
				}
				bool first = true;
				for(uint32_t i = 0; i < nmap; i++) {
					if(parent == map[i].taxID) {
						if(!first) { if(i + 1 < nmap) map[i] = map[nmap - 1]; nmap--; i--; }
						else first = false;
					}
				}
				if(nmap <= k) break;
			}
			++rank;
			if(rank > path_size(map[0])) break;
		}
	}
	if(!only_host && nmap > k) return 0;
	uint32_t no = 0;
	for(uint32_t i = 0; i < nmap; i++) {
		if(only_host && !is_host(v, map[i].taxID)) continue;
		OutRec r; r.taxid = map[i].taxID; r.score = map[i].score; r.hitlen = map[i].hitlen;
		r.uid = map[i].uniqueID < (uint64_t)v.n_seqs ? (uint32_t)map[i].uniqueID : 0xFFFFFFFFu; r.pad = 0;
		out[no++] = r;
	}
	return no;
}

// Scalar resolve of one SA row (walk-left to a sampled / boundary / '$' row; no "+steps",
// group_walk.h:508-512; tryOffset bt2_idx.h:1980-2014).  Scalar twin of the cooperative kernel.
CFB_HDN uint32_t resolve_scalar(const IndexView& v, uint64_t row, Counters* ctr) {
	const uint64_t lowmask = ((uint64_t)1 << v.off_rate) - 1;
	for(;;) {
		if(row == v.zoff) return 0;
		if((row & lowmask) == 0) {
			const uint64_t i = row >> v.off_rate;
			return v.sample32 ? v.sample32[i] : (uint32_t)v.sample16[i];
		}
		if(v.n_boundaries && row <= v.last_boundary) {
			const uint64_t b = row >> v.bshift;
			if(v.bbits[b >> 5] & (1u << (b & 31))) {
				uint32_t lo = 0, hi = v.n_boundaries;
				while(lo < hi) { uint32_t mid = (lo + hi) >> 1; if(v.brow[mid] < row) lo = mid + 1; else hi = mid; }
				if(lo < v.n_boundaries && v.brow[lo] == row) return v.sample32 ? v.bseq[lo] : (uint32_t)(uint16_t)v.bseq[lo];
			}
		}
		row = lf_scalar(v, row, bwt_char(v, row));
		if(ctr) ctr->walk_steps++;
	}
}

}  // namespace cfb
#endif
