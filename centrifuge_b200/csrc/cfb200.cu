// cfb200.cu -- kernels + C ABI (include/cfb200.h) of the B200-native classification path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see centrifuge_b200/build.py).
#include "../../include/cfb200.h"
#include "cf_index.h"
#include "cf_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <fcntl.h>
#include <unistd.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

using namespace cfb;

// =======================================================================================
// k_search
// =======================================================================================
enum { M_DONE = 0, M_FTAB = 1, M_LF = 2, M_NEED = 3, M_FTABK = 4, M_FTABD = 5 };

struct Walk {            // group-uniform state of one greedy strand walk
	const uint8_t* fw; uint32_t rlen; int strand; uint32_t tid;
	uint32_t cur, dep, offset, nh;
	uint64_t top, bot, fi;
	uint32_t tnext, tend;
	int mode;
};

struct SearchArgs {
	IndexView v; Params p; BatchView b;
	HitRec* hits; uint32_t* nhits; uint32_t cap;
	unsigned int* task_ctr; unsigned long long* task_ctr64; uint32_t ntasks, chunk;
	unsigned int* overflow;
	Counters* ctr;
	const uint64_t* pk; const uint32_t* nm; uint32_t W;   // packed reads (k_pack)
	uint32_t jump_w;                                      // widest range advanced eight bases per gather through walk8 (CFB_JUMP_W, default 4)
	uint32_t keep_short;                                  // store every hit (min_hitlen < 22, generic kernels); else only hits of >= kLongLen bases
};

// row -> (side, offset in side).  Rows are < 2^39 for any index that fits in HBM, so row>>7 fits
// 32 bits and the division by 3 is a single mul.hi (384 = 128 * 3).
__device__ __forceinline__ void row_locus(uint64_t row, uint64_t& side, uint32_t& off) {
	const uint32_t q = (uint32_t)(row >> 7) / 3u;
	side = q; off = (uint32_t)(row - (uint64_t)q * 384u);
}
__device__ __forceinline__ uint64_t shl64(uint64_t v, uint32_t n) {   // PTX shl clamps n >= 64 to "all shifted out"
	uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(n)); return r;
}
__device__ __forceinline__ uint64_t shr64(uint64_t v, uint32_t n) {
	uint64_t r; asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(v), "r"(n)); return r;
}

// =======================================================================================
// k_pack: reads -> 2-bit words in *consumption order* of the backward search, both strands.
// Position p of a packed strand is the base the search looks at when dep == p:
//   strand 0 (fw): fw[rlen-1-p]      strand 1 (rc): comp(fw[p])      (Read::constructRevComps)
// base p sits in bits 2*(p&31) of word p>>5; N -> code 0 plus a bit in the parallel N mask.
// =======================================================================================
struct PackArgs { BatchView b; uint64_t* pk; uint32_t* nm; uint32_t W; };

// 4 bytes at an arbitrary byte offset of a 4-byte aligned buffer (reads at most 7 bytes past `off`: device
// buffers carry that much slack)
__device__ __forceinline__ uint32_t load4(const uint8_t* base, uint64_t off) {
	const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (off & ~3ull));
	return __funnelshift_r(w[0], w[1], (uint32_t)(off & 3) * 8);
}

// four 2-bit fields held one per byte -> 8 contiguous bits; four 1-bit flags held one per byte -> 4 bits
__device__ __forceinline__ uint32_t squeeze4x2(uint32_t x) { x = (x | (x >> 6)) & 0x000F000Fu; return (x | (x >> 12)) & 0xFFu; }
__device__ __forceinline__ uint32_t squeeze4x1(uint32_t x) { x = (x | (x >> 7)) & 0x00030003u; return (x | (x >> 14)) & 0xFu; }

__global__ void __launch_bounds__(128) k_pack(const PackArgs a) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (task, word)
	const uint64_t ntasks = (uint64_t)a.b.n_units * a.b.n_mates * 2;
	if(i >= ntasks * a.W) return;
	const uint64_t t = i / a.W; const uint32_t k = (uint32_t)(i - t * a.W);
	const uint32_t per = 2u * (uint32_t)a.b.n_mates;
	const uint32_t unit = (uint32_t)(t / per), rem = (uint32_t)(t - (uint64_t)unit * per);
	const int mate = (int)(rem >> 1), strand = (int)(rem & 1);
	const uint32_t len = mate ? a.b.len[1][unit] : a.b.len[0][unit];        // no runtime index into the parameter arrays
	uint64_t w = 0; uint32_t n = 0;
	if(k * 32 < len) {
		const uint64_t off = mate ? a.b.off[1][unit] : a.b.off[0][unit];
		const uint32_t cnt = min(32u, len - k * 32);         // bases of this word
		// strand 1 consumes fw[p], strand 0 consumes fw[len-1-p]: either way a window of `cnt` consecutive bytes,
		// packed in window order four bytes at a time and flipped end to end for strand 0
		const uint64_t lo = strand ? off + k * 32 : off + (len - k * 32 - cnt);
		#pragma unroll
		for(uint32_t g = 0; g < 32; g += 4) {
			if(g < cnt) {
				const uint32_t keep = cnt - g >= 4 ? 0xffffffffu : ((1u << (8 * (cnt - g))) - 1u);
				const uint32_t v = load4(a.b.bases, lo + g) & keep;
				const uint32_t nb = __vcmpne4(v & 0xFCFCFCFCu, 0u) & 0x01010101u;        // codes above 3 are N
				uint32_t b2 = v & 0x03030303u;
				if(strand) b2 ^= 0x03030303u & keep;
				b2 &= ~(nb * 3u);
				w |= (uint64_t)squeeze4x2(b2) << (2 * g);
				n |= squeeze4x1(nb) << g;
			}
		}
		if(!strand) {
			uint64_t x = __brevll(w);
			x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
			w = x >> (2 * (32 - cnt)); n = __brev(n) >> (32 - cnt);
		}
	}
	a.pk[i] = w; a.nm[i] = n;
}

// G lanes own one walk; lane gl holds the NL = 8/G 16-byte pieces q = gl*NL + j of a 128-byte side
// (q < 6: 64 BWT bases each, q = 6: occ[A],occ[C], q = 7: occ[G],occ[T]).
template <int G> struct Side { uint4 d[8 / G]; };

template <int G>
__device__ __forceinline__ void side_load(Side<G>& r, const uint4* sides4, uint64_t s, unsigned gl, uint32_t need_bases, int c) {
	constexpr int NL = 8 / G;
	const uint4* p = sides4 + s * 8 + gl * NL;
	#pragma unroll
	for(int j = 0; j < NL; j++) {
		const int q = (int)gl * NL + j;
		// only the sectors the rank needs: BWT pieces below `need_bases`, and the occ pair of base c
		const bool want = q < 6 ? ((uint32_t)(64 * q) < need_bases) : (q == 6 + (c >> 1));
		r.d[j] = want ? __ldg(p + j) : make_uint4(0, 0, 0, 0);
	}
}
// match mask of one 16-byte piece as two u64 (bit 2i set <=> base i == c)
__device__ __forceinline__ void piece_match(const uint4& d, uint64_t rep, uint64_t& m0, uint64_t& m1) {
	const uint64_t v0 = (uint64_t)d.x | ((uint64_t)d.y << 32), v1 = (uint64_t)d.z | ((uint64_t)d.w << 32);
	const uint64_t x0 = ~(v0 ^ rep), x1 = ~(v1 ^ rep);
	m0 = x0 & (x0 >> 1) & 0x5555555555555555ull; m1 = x1 & (x1 >> 1) & 0x5555555555555555ull;
}
// number of set match bits among the first n (0..64) bases of a piece
__device__ __forceinline__ uint32_t piece_prefix(uint64_t m0, uint64_t m1, int n) {
	const int n0 = n < 32 ? n : 32, n1 = n - 32;            // n1 may be negative
	uint32_t r = __popcll(shl64(m0, (uint32_t)(64 - 2 * n0)));
	if(n1 > 0) r += __popcll(shl64(m1, (uint32_t)(64 - 2 * n1)));
	return r;
}
// rank of base c below offT and offB in the same side; returns cT | cB << 16 (this lane's share)
template <int G>
__device__ __forceinline__ uint32_t side_count2(const Side<G>& r, unsigned gl, uint64_t rep, uint32_t offT, uint32_t offB) {
	constexpr int NL = 8 / G;
	uint32_t acc = 0;
	#pragma unroll
	for(int j = 0; j < NL; j++) {
		const int q = (int)gl * NL + j;
		if(q < 6) {
			uint64_t m0, m1; piece_match(r.d[j], rep, m0, m1);
			int kT = (int)offT - 64 * q; kT = kT < 0 ? 0 : (kT > 64 ? 64 : kT);
			int kB = (int)offB - 64 * q; kB = kB < 0 ? 0 : (kB > 64 ? 64 : kB);
			acc += piece_prefix(m0, m1, kT) | (piece_prefix(m0, m1, kB) << 16);
		}
	}
	return acc;
}
template <int G>
__device__ __forceinline__ uint32_t side_count1(const Side<G>& r, unsigned gl, uint64_t rep, uint32_t off) {
	constexpr int NL = 8 / G;
	uint32_t acc = 0;
	#pragma unroll
	for(int j = 0; j < NL; j++) {
		const int q = (int)gl * NL + j;
		if(q < 6) {
			uint64_t m0, m1; piece_match(r.d[j], rep, m0, m1);
			int k = (int)off - 64 * q; k = k < 0 ? 0 : (k > 64 ? 64 : k);
			acc += piece_prefix(m0, m1, k);
		}
	}
	return acc;
}
template <int G>
__device__ __forceinline__ uint32_t gsum(uint32_t x, unsigned gmask) {
	#pragma unroll
	for(int o = 1; o < G; o <<= 1) x += __shfl_xor_sync(gmask, x, o);
	return x;
}
template <int G>
__device__ __forceinline__ uint64_t side_occ(const Side<G>& r, int c, unsigned gmask, unsigned gbase) {
	constexpr int NL = 8 / G;
	const int q = 6 + (c >> 1), j = q % NL;
	uint4 v = r.d[NL - 1];
	#pragma unroll
	for(int jj = 0; jj < NL - 1; jj++) if(jj == j) v = r.d[jj];
	uint32_t lo = (c & 1) ? v.z : v.x, hi = (c & 1) ? v.w : v.y;
	if(G > 1) { const int src = gbase + q / NL; lo = __shfl_sync(gmask, lo, src); hi = __shfl_sync(gmask, hi, src); }
	return (uint64_t)lo | ((uint64_t)hi << 32);
}
template <int G>
__device__ __forceinline__ int side_char(const Side<G>& r, uint32_t off, unsigned gmask, unsigned gbase) {
	constexpr int NL = 8 / G;
	const int q = (int)(off >> 6), j = q % NL;
	uint4 v = r.d[0];
	#pragma unroll
	for(int jj = 1; jj < NL; jj++) if(jj == j) v = r.d[jj];
	const uint32_t w = sel_word(v, (off >> 4) & 3);
	uint32_t ch = (w >> ((off & 15) * 2)) & 3;
	if(G > 1) ch = __shfl_sync(gmask, ch, gbase + q / NL);
	return (int)ch;
}

struct Walk2 {           // group-uniform state of one greedy strand walk over a packed read
	const uint64_t* pk; const uint32_t* nm;   // packed strand of the current task
	uint32_t rlen, tid, cur, dep, offset, nh, tnext, tend;
	uint64_t top, bot, fi;
	uint64_t rw0, rw1; uint32_t nw0, nw1, rwi;   // packed words rwi, rwi+1 and their N masks
	int mode;
};

template <bool COUNT, int G>
struct SearchCtx {
	const SearchArgs& a; unsigned gmask, gbase, gl;
	unsigned long long c_ps, c_ft, c_sides, c_lf;
	bool pooled;        // thread-per-walk kernels: tasks are handed out at the loop top (pool_take)
	__device__ __forceinline__ SearchCtx(const SearchArgs& a_) : a(a_), c_ps(0), c_ft(0), c_sides(0), c_lf(0), pooled(false) {
		const unsigned lane = threadIdx.x & 31;
		gl = lane & (G - 1); gbase = lane - gl; gmask = ((G == 32) ? 0xffffffffu : ((1u << G) - 1u)) << gbase;
	}
	__device__ __forceinline__ void emit(Walk2& w, uint64_t top, uint64_t bot, uint32_t off, uint32_t len) {
		if(w.nh < a.cap) {
			if(gl == 0) { HitRec* h = a.hits + (size_t)w.tid * a.cap + w.nh; h->top = top; h->bot = bot; h->bwoff = off; h->len = len; }
		} else if(gl == 0) atomicExch(a.overflow, 1u);
		w.nh++;
	}
	__device__ __forceinline__ void load_words(Walk2& w, uint32_t wi) {
		w.rw0 = __ldg(w.pk + wi); w.rw1 = __ldg(w.pk + wi + 1);
		w.nw0 = __ldg(w.nm + wi); w.nw1 = __ldg(w.nm + wi + 1);
		w.rwi = wi;
	}
	// bind walk state to task id w.tid; false if the mate is filtered / empty
	__device__ __forceinline__ bool bind_task(Walk2& w) {
		const uint32_t per = 2u * (uint32_t)a.b.n_mates;
		const uint32_t unit = w.tid / per, rem = w.tid - unit * per;
		const int mate = (int)(rem >> 1);
		const uint8_t fl = a.b.flags ? a.b.flags[unit] : 3;
		w.nh = 0;
		w.rlen = a.b.len[mate][unit];
		if(!((fl >> mate) & 1) || w.rlen == 0) { a.nhits[w.tid] = 0; return false; }
		w.pk = a.pk + (size_t)w.tid * a.W; w.nm = a.nm + (size_t)w.tid * a.W;
		w.cur = 0;
		return true;
	}
	__device__ __forceinline__ bool next_task(Walk2& w) {
		if(pooled) { w.mode = M_NEED; return false; }
		for(;;) {
			if(w.tnext >= w.tend) {
				unsigned base = 0;
				if(gl == 0) base = atomicAdd(a.task_ctr, a.chunk);
				if(G > 1) base = __shfl_sync(gmask, base, gbase);
				if(base >= a.ntasks) { w.mode = M_DONE; return false; }
				w.tnext = base; w.tend = min(base + a.chunk, a.ntasks);
			}
			w.tid = w.tnext++;
			const uint32_t per = 2u * (uint32_t)a.b.n_mates;
			const uint32_t unit = w.tid / per, rem = w.tid - unit * per;
			const int mate = (int)(rem >> 1);
			const uint8_t fl = a.b.flags ? a.b.flags[unit] : 3;
			w.nh = 0;
			w.rlen = a.b.len[mate][unit];
			if(!((fl >> mate) & 1) || w.rlen == 0) { if(gl == 0) a.nhits[w.tid] = 0; continue; }
			w.pk = a.pk + (size_t)w.tid * a.W; w.nm = a.nm + (size_t)w.tid * a.W;
			w.cur = 0;
			return true;
		}
	}
	__device__ __forceinline__ void finish_task(Walk2& w) { if(gl == 0) a.nhits[w.tid] = (w.nh & 0x7fffu) | ((w.nh < 0x7fffu ? w.nh : 0x7fffu) << 15); }   // every hit stored; word format of nh_pack (defined further down)

	// Starts partial searches at w.cur until one needs the ftab (mode M_FTAB) or work runs out.
	__device__ __forceinline__ void start_search(Walk2& w) {
		const uint32_t fc = (uint32_t)a.v.ftab_chars;
		for(;;) {
			if(COUNT) c_ps++;
			w.offset = w.cur;
			const uint32_t left = w.rlen - w.cur;
			if(left < fc) {                               // hi_aligner.h:939-949
				emit(w, kOff, kOff, w.offset, w.rlen - w.offset);
				finish_task(w);
				if(!next_task(w)) return;
				continue;
			}
			load_words(w, w.cur >> 5);
			const uint32_t sh = w.cur & 31;
			const uint64_t win = shr64(w.rw0, 2 * sh) | shl64(w.rw1, 64 - 2 * sh);      // bases cur.., base cur in the low bits
			const uint32_t nwin = (uint32_t)(((uint64_t)w.nw0 | ((uint64_t)w.nw1 << 32)) >> sh);
			const uint32_t nbits = nwin & ((1u << fc) - 1u);
			if(nbits) {                                   // N within the next fc bases, hi_aligner.h:951-966
				const uint32_t hl = (uint32_t)__ffs(nbits);  // index of the first N + 1
				w.cur += hl;
				emit(w, kOff, kOff, w.offset, hl);
				if(after_hit(w, hl)) continue; else return;
			}
			w.fi = win & ((1ull << (2 * fc)) - 1ull);     // base cur (rightmost of the 10-mer) least significant
			w.mode = M_FTAB;
			return;
		}
	}
	// searchForwardAndReverse restart policy (classifier.h:686-766).
	__device__ __forceinline__ bool after_hit(Walk2& w, uint32_t hlen) {
		bool done = w.cur >= w.rlen;
		if(!done) {
			if(hlen > a.p.increment) w.cur += 1;
			if(w.cur + a.p.min_hitlen >= w.rlen) done = true;
		}
		if(done) { finish_task(w); if(!next_task(w)) return false; }
		return true;
	}
	__device__ __forceinline__ void hit_and_restart(Walk2& w) {
		const uint32_t hl = w.dep - w.offset;
		emit(w, w.top, w.bot, w.offset, hl);
		w.cur = w.dep;
		if(after_hit(w, hl)) start_search(w);
	}
};

template <bool COUNT, int G>
__global__ void __launch_bounds__(kSearchThreads) k_search(const SearchArgs a) {
	SearchCtx<COUNT, G> cx(a);
	const unsigned gl = cx.gl, gmask = cx.gmask, gbase = cx.gbase;
	const uint4* sides4 = reinterpret_cast<const uint4*>(a.v.sides);
	Walk2 w; memset(&w, 0, sizeof w); w.mode = M_DONE;
	if(cx.next_task(w)) cx.start_search(w);

	while(__any_sync(0xffffffffu, w.mode != M_DONE)) {
		// ---------------- single fetch point ----------------
		uint32_t offT = 0, offB = 0; uint64_t sT = 0, sB = 0; bool range = false, same = true;
		int c = 4;
		uint64_t e0 = 0, e1 = 0;
		Side<G> da, db;
		if(w.mode == M_FTAB) {
			if(G == 1) { e0 = __ldg(a.v.ftab + w.fi); e1 = __ldg(a.v.ftab + w.fi + 1); }
			else if(gl < 2) e0 = __ldg(a.v.ftab + w.fi + gl);
		} else if(w.mode == M_LF) {
			const uint32_t wi = w.dep >> 5;
			if(wi != w.rwi && wi != w.rwi + 1) cx.load_words(w, wi);
			const uint32_t sh = w.dep & 31;
			const uint64_t rw = wi == w.rwi ? w.rw0 : w.rw1; const uint32_t nw = wi == w.rwi ? w.nw0 : w.nw1;
			c = ((nw >> sh) & 1u) ? 4 : (int)((rw >> (2 * sh)) & 3);
			if(c <= 3) {
				row_locus(w.top, sT, offT);
				const uint64_t spread = w.bot - w.top;
				range = spread != 1;
				sB = sT; offB = offT;
				if(range) {
					same = spread < (uint64_t)(384 - offT);   // initFromTopBot bt2_idx.h:326-349
					if(same) offB = offT + (uint32_t)spread; else row_locus(w.bot, sB, offB);
				}
				// rank needs bases [0, off); mapLF1 additionally needs BWT[offT]
				side_load<G>(da, sides4, sT, gl, same ? (range ? offB : offT + 1) : offT, c);
				if(!same) side_load<G>(db, sides4, sB, gl, offB, c);
			}
		}
		// ---------------- consume ----------------
		if(w.mode == M_FTAB) {
			if(G > 1) { const uint64_t mine = e0; e0 = __shfl_sync(gmask, mine, gbase); e1 = __shfl_sync(gmask, mine, gbase + 1); }
			if(COUNT) cx.c_ft++;
			w.top = ftab_hi(a.v, e0); w.bot = ftab_lo(a.v, e1);
			w.dep = w.cur + (uint32_t)a.v.ftab_chars;
			if(w.bot <= w.top) {                          // hi_aligner.h:971-982
				const uint32_t hl = w.dep - w.offset;
				cx.emit(w, kOff, kOff, w.offset, hl);
				w.cur = w.dep;
				if(cx.after_hit(w, hl)) cx.start_search(w);
			} else if(w.dep < w.rlen) w.mode = M_LF;
			else cx.hit_and_restart(w);
		} else if(w.mode == M_LF) {
			bool fail = c > 3;
			uint64_t t = 0, b = 0;
			if(!fail) {
				const uint64_t rep = (uint64_t)c * 0x5555555555555555ull;
				uint32_t packed;
				if(same) packed = side_count2<G>(da, gl, rep, offT, offB);
				else packed = side_count1<G>(da, gl, rep, offT) | (side_count1<G>(db, gl, rep, offB) << 16);
				packed = gsum<G>(packed, gmask);
				const uint64_t occT = side_occ<G>(da, c, gmask, gbase);
				const uint64_t occB = same ? occT : side_occ<G>(db, c, gmask, gbase);
				uint64_t rT = packed & 0xFFFFu, rB = packed >> 16;
				if(c == 0) {
					if(sT == a.v.zside && a.v.zoffc < offT) rT--;
					if(sB == a.v.zside && a.v.zoffc < offB) rB--;
				}
				t = a.v.fchr[c] + occT + rT;
				if(range) {
					b = a.v.fchr[c] + occB + rB;
					if(COUNT) { cx.c_lf += 2; cx.c_sides += same ? 1 : 2; }
				} else {                                  // mapLF1 bt2_idx.h:2910-2933
					if(COUNT) { cx.c_lf += 1; cx.c_sides += 1; }
					const int rowc = side_char<G>(da, offT, gmask, gbase);
					if(rowc != c || w.top == a.v.zoff) fail = true;
					b = t + 1;
				}
				if(b <= t) fail = true;
			}
			if(fail) cx.hit_and_restart(w);
			else {
				w.top = t; w.bot = b; w.dep++;
				if(w.dep >= w.rlen) cx.hit_and_restart(w);
			}
		}
	}
	if(COUNT && gl == 0 && a.ctr) {
		atomicAdd(&a.ctr->partial_searches, cx.c_ps); atomicAdd(&a.ctr->ftab_probes, cx.c_ft);
		atomicAdd(&a.ctr->sides_search, cx.c_sides); atomicAdd(&a.ctr->lf_steps, cx.c_lf);
	}
}

// ---------------------------------------------------------------------------------------
// Work distribution for the thread-per-walk kernels: a warp owns a pool [base, end) of task ids
// refilled with ONE global atomic per `chunk` (>= 32) tasks; lanes that need work take consecutive
// ids by ballot rank at a convergent point of the loop.  (A per-lane atomicAdd on one address
// serialises in L2: ~1M same-address atomics cost more than the whole resolve kernel.)
// ---------------------------------------------------------------------------------------
struct WarpPool { unsigned long long base, end; };
__device__ __forceinline__ bool pool_take(WarpPool& P, bool want, unsigned long long* ctr, unsigned long long total, unsigned chunk, unsigned long long& out) {
	const unsigned need = __ballot_sync(0xffffffffu, want);
	if(!need) return false;
	const unsigned lane = threadIdx.x & 31;
	const unsigned cnt = __popc(need), r = __popc(need & ((1u << lane) - 1u));
	const unsigned long long avail = P.end - P.base;
	unsigned long long nb = 0;
	const bool refill = avail < cnt;
	if(refill) { if(lane == 0) nb = atomicAdd(ctr, (unsigned long long)chunk); nb = __shfl_sync(0xffffffffu, nb, 0); }
	unsigned long long t, lim;
	if(r < avail) { t = P.base + r; lim = P.end; }
	else { t = nb + (r - avail); lim = nb + chunk < total ? nb + chunk : total; }
	if(refill) { P.base = nb + (cnt - avail); P.end = nb + chunk < total ? nb + chunk : total; if(P.base > P.end) P.base = P.end; }
	else P.base += cnt;
	out = t;
	return want && t < lim;
}

// ---------------------------------------------------------------------------------------
// Re-blocked device index.  The `.cf` side (96 B of BWT = 384 rows + 4 x u64 occ) is the unit the
// reference's CPU code walks; for the GPU each side is split at load time into three 64-byte
// blocks of 128 rows:  u64 occ[A,C,G,T] counted before the block ('$' excluded) | 4 x u64 of BWT.
// One LF step then touches half a cache line (two 32-byte sectors), the row -> block map is a shift,
// and the rank runs over at most 128 bases.  Results are identical to countBt2Side (bt2_idx.h:2192).
// ---------------------------------------------------------------------------------------
__global__ void k_build_blocks(const uint64_t* sides, uint64_t num_sides, uint64_t zside, uint32_t zoffc, uint64_t* blocks) {
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(b > num_sides * 3) return;
	// block num_sides*3 is a sentinel holding the total counts: an exclusive bound bot == len+1 lands
	// on it when len+1 is a multiple of 384 (then the last side has no padding and the totals are exact)
	const bool sentinel = b == num_sides * 3;
	const uint64_t s = sentinel ? num_sides - 1 : b / 3; const uint32_t part = sentinel ? 3u : (uint32_t)(b - s * 3);
	const uint64_t* sd = sides + s * 16;
	uint64_t occ[4] = {sd[12], sd[13], sd[14], sd[15]};
	for(uint32_t k = 0; k < part * 4; k++) {
		const uint64_t w = sd[k];
		const uint64_t lo = w & 0x5555555555555555ull, hi = (w >> 1) & 0x5555555555555555ull;
		const uint32_t c1 = __popcll(lo & ~hi), c2 = __popcll(hi & ~lo), c3 = __popcll(hi & lo);
		occ[1] += c1; occ[2] += c2; occ[3] += c3; occ[0] += 32 - c1 - c2 - c3;
	}
	if(s == zside && zoffc < part * 128) occ[0] -= 1;        // the '$' row is stored as A but is not an A
	uint64_t* o = blocks + b * 8;
	o[0] = occ[0]; o[1] = occ[1]; o[2] = occ[2]; o[3] = occ[3];
	for(uint32_t k = 0; k < 4; k++) o[4 + k] = sentinel ? 0ull : sd[part * 4 + k];
}

// ---------------------------------------------------------------------------------------
// Per-base rank sectors (the search kernel's view of the index).  For every 192 rows (half a side)
// and every base c one 32-byte sector:  u64 occ_c before the block | 192 indicator bits BWT[row]==c.
// LF(row, c) = fchr[c] + occ + popcount(bits below row) touches exactly one 32-byte DRAM sector;
// the four sectors of a block share one 128-byte line.  The '$' row's bit is cleared in A's vector,
// so neither the rank nor the mapLF1 test needs a '$' special case.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t even_bits(uint64_t x) {    // gather bits 0,2,4,.. into the low 32 bits
	x &= 0x5555555555555555ull;
	x = (x | (x >> 1)) & 0x3333333333333333ull;
	x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
	x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
	x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
	x = (x | (x >> 16)) & 0x00000000ffffffffull;
	return x;
}
__global__ void k_build_rankv(const uint64_t* sides, uint64_t num_sides, uint64_t zside, uint32_t zoffc, uint64_t* rankv) {
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(b > num_sides * 2) return;
	const bool sentinel = b == num_sides * 2;                 // totals, for an exclusive bound == len+1 on a side boundary
	const uint64_t s = sentinel ? num_sides - 1 : b >> 1; const uint32_t half = sentinel ? 2u : (uint32_t)(b & 1);
	const uint64_t* sd = sides + s * 16;
	uint64_t occ[4] = {sd[12], sd[13], sd[14], sd[15]};
	for(uint32_t k = 0; k < half * 6; k++) {
		const uint64_t w = sd[k];
		const uint64_t lo = w & 0x5555555555555555ull, hi = (w >> 1) & 0x5555555555555555ull;
		const uint32_t c1 = __popcll(lo & ~hi), c2 = __popcll(hi & ~lo), c3 = __popcll(hi & lo);
		occ[1] += c1; occ[2] += c2; occ[3] += c3; occ[0] += 32 - c1 - c2 - c3;
	}
	if(s == zside && zoffc < half * 192) occ[0] -= 1;
	for(int c = 0; c < 4; c++) {
		uint64_t bits[3] = {0, 0, 0};
		if(!sentinel) {
			for(int k = 0; k < 6; k++) {
				const uint64_t m = even_bits(match2(sd[half * 6 + k], c));       // 32 indicator bits
				bits[k >> 1] |= m << (32 * (k & 1));
			}
			if(c == 0 && s == zside && zoffc >= half * 192 && zoffc < half * 192 + 192) { const uint32_t z = zoffc - half * 192; bits[z >> 6] &= ~(1ull << (z & 63)); }
		}
		uint64_t* o = rankv + (b * 4 + c) * 4;
		o[0] = occ[c]; o[1] = bits[0]; o[2] = bits[1]; o[3] = bits[2];
	}
}
// rows below `off` (0..191) whose indicator bit is set
__device__ __forceinline__ uint32_t rankv_count(uint64_t b0, uint64_t b1, uint64_t b2, uint32_t off) {
	const uint32_t k = off >> 6, part = off & 63;
	const uint64_t bk = k == 0 ? b0 : (k == 1 ? b1 : b2);
	uint32_t r = (uint32_t)__popcll(bk & ((1ull << part) - 1ull));
	if(k > 0) r += (uint32_t)__popcll(b0);
	if(k > 1) r += (uint32_t)__popcll(b1);
	return r;
}
__device__ __forceinline__ void row_block192(uint64_t row, uint64_t& blk, uint32_t& off) {
	blk = __umul64hi(row >> 6, 0xAAAAAAAAAAAAAAABull) >> 1;     // (row / 64) / 3
	off = (uint32_t)(row - blk * 192);
}

// ---------------------------------------------------------------------------------------
// rank16: the layout both walk kernels use.  For every 64 rows and every base c one 16-byte entry
//   u64 occ_c (count of c before the block, '$' excluded; bit 63 of A's entry = "block holds a genome-
//   boundary row")  |  u64 indicator bits (BWT[row] == c; the '$' row has no bit)
// so LF(row, c) = fchr[c] + occ + popc(bits & lowmask(row & 63)) costs ONE 16-byte load request.
// Measured on this part (tools/gather_bench.cu): fully divergent gathers are capped at ~70 G 16-byte
// lane requests/s independent of size (32 B: 34 G/s, 64 B: 17.6 G/s, 128 B: 8.9 G/s), so requests per
// LF step -- not bytes -- is what bounds the walk.  The four bases of a block share one 64-byte chunk.
// ---------------------------------------------------------------------------------------
__global__ void k_build_rank16(const uint64_t* sides, uint64_t num_sides, uint64_t zside, uint32_t zoffc, uint64_t* r16) {
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(b > num_sides * 6) return;
	const bool sentinel = b == num_sides * 6;                 // totals, for an exclusive bound == len+1 on a side boundary
	const uint64_t s = sentinel ? num_sides - 1 : b / 6; const uint32_t sub = sentinel ? 6u : (uint32_t)(b - s * 6);
	const uint64_t* sd = sides + s * 16;
	uint64_t occ[4] = {sd[12], sd[13], sd[14], sd[15]};
	for(uint32_t k = 0; k < sub * 2; k++) {
		const uint64_t w = sd[k];
		const uint64_t lo = w & 0x5555555555555555ull, hi = (w >> 1) & 0x5555555555555555ull;
		const uint32_t c1 = __popcll(lo & ~hi), c2 = __popcll(hi & ~lo), c3 = __popcll(hi & lo);
		occ[1] += c1; occ[2] += c2; occ[3] += c3; occ[0] += 32 - c1 - c2 - c3;
	}
	if(s == zside && zoffc < sub * 64) occ[0] -= 1;
	for(int c = 0; c < 4; c++) {
		uint64_t bits = 0;
		if(!sentinel) {
			bits = even_bits(match2(sd[sub * 2], c)) | (even_bits(match2(sd[sub * 2 + 1], c)) << 32);
			if(c == 0 && s == zside && zoffc >= sub * 64 && zoffc < sub * 64 + 64) bits &= ~(1ull << (zoffc - sub * 64));
		}
		r16[(b * 4 + c) * 2] = occ[c]; r16[(b * 4 + c) * 2 + 1] = bits;
	}
}
__global__ void k_mark_boundaries(const uint64_t* brow, uint32_t n, uint64_t* r16) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) atomicOr((unsigned long long*)&r16[(brow[i] >> 6) * 8], 1ull << 63);
}
__global__ void k_build_ftab2(IndexView v, uint64_t n, uint64_t* ftab2) {
	const uint64_t fi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(fi >= n) return;
	ftab2[fi * 2] = ftab_hi(v, v.ftab[fi]); ftab2[fi * 2 + 1] = ftab_lo(v, v.ftab[fi + 1]);
}
// nhits[] word of a strand list: hits stored (15 bits) | hits found (15 bits, saturating) << 15 | kListNoLong.  With
// min_hitlen >= 22 the search kernel stores only hits of at least kLongLen bases: shorter ones are never counted
// (classifier.h:299), sort after every stored hit (compareBWTHits puts len >= 22 first) and touch nothing else -- unless
// both strands of a mate are in play (extension / twin removal, classifier.h:790-870) or a list is long enough for
// introsort (> 16 hits, where libstdc++'s tie permutation may depend on every element).  In those rare cases k_prep
// regenerates the full lists with a one-thread version of the kernel's walk (search_strand_dev) into a side buffer and points
// the list at it: word = kListRegen | slot.  Nine of ten hits of a typical read are short, so this removes most of the hit
// traffic (random partial-sector writes) and shrinks the per-read device footprint from ~3.6 KB to ~2.3 KB.
static const uint32_t kListNoLong = 0x80000000u;             // the strand has no hit of min_hitlen bases
static const uint32_t kListRegen = 0x40000000u;              // the list lives in the regeneration buffer, slot = low 30 bits (written by k_prep)
static const uint32_t kLongLen = 22;
__host__ __device__ __forceinline__ uint32_t nh_pack(uint32_t stored, uint32_t found, bool nolong) {
	return (stored & 0x7fffu) | ((found < 0x7fffu ? found : 0x7fffu) << 15) | (nolong ? kListNoLong : 0u);
}
__host__ __device__ __forceinline__ uint32_t nh_stored(uint32_t w) { return w & 0x7fffu; }
__host__ __device__ __forceinline__ uint32_t nh_found(uint32_t w) { return (w >> 15) & 0x7fffu; }
static const uint64_t kOccMask = 0x7fffffffffffffffull;
static const uint64_t kWalkRowMask = (1ull << 40) - 1ull;   // walk8 entry: row in the low 40 bits
static const int kJumpRows __attribute__((unused)) = 1;     // widest range advanced through walk8 (the kernel is specialised for 1).  Ranges of 2-4 adjacent rows can take the same
                                                            // jump (LF keeps them adjacent), but on the bench workload they shrink so often that the
                                                            // failed attempts cost more than the jumps save: 4.44 ms vs 4.25 ms with 1 (measured)

// Extended jump table: the SA range of every K-mer (K > ftabChars), obtained by K - ftabChars LF steps from the
// 10-mer range -- exactly what partialSearch would compute base by base (hi_aligner.h:985-1008).  It trades
// HBM capacity (16 B x 4^K; 17 GB at K = 15) for random accesses: one gather replaces K - 10 walk steps whose
// top and bot rows are far apart (two DRAM sectors each).  An empty entry (bot <= top) only says the range
// died somewhere in between; the kernel then redoes that search from the 10-mer table to find where.
__global__ void k_build_ftabk(IndexView v, int K, uint64_t n, uint64_t* out) {
	const uint64_t fk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(fk >= n) return;
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(v.rank16);
	const int fc = v.ftab_chars;
	const uint64_t f10 = fk & ((1ull << (2 * fc)) - 1ull);
	uint64_t top = v.ftab2[f10 * 2], bot = v.ftab2[f10 * 2 + 1];
	for(int j = fc; j < K && bot > top; j++) {
		const int c = (int)((fk >> (2 * j)) & 3);
		const ulonglong2 tq = __ldg(r16 + (top >> 6) * 4 + c), bq = __ldg(r16 + (bot >> 6) * 4 + c);
		top = v.fchr[c] + (tq.x & kOccMask) + (uint64_t)__popcll(tq.y & ((1ull << (top & 63)) - 1ull));
		bot = v.fchr[c] + (bq.x & kOccMask) + (uint64_t)__popcll(bq.y & ((1ull << (bot & 63)) - 1ull));
	}
	if(bot <= top) { top = 0; bot = 0; }
	out[fk * 2] = top; out[fk * 2 + 1] = bot;
}

// Death-depth table: for every (K+3)-mer, how far a partial search that starts with it gets before its SA range empties,
// as far as that can be said in 2 bits:  1, 2, 3 = the K-mer occurs and the range dies after K, K+1, K+2 bases (the hit
// partialSearch would report has exactly that length);  0 = the (K+3)-mer occurs, or the K-mer does not (go through the jump
// table).  A partial search on a strand that does not match -- most searches of most reads -- dies within these three bases
// and then costs ONE gather instead of the K-mer lookup plus one or two rank gathers per base; its SA range is not
// computed, which is fine because a hit shorter than min_hitlen is never resolved (k_prep recomputes the range in the
// rare cases where a short hit's size can influence the result).  Layout: K-mer major, 64 extensions = 16 bytes per K-mer;
// extension e = c_K | c_{K+1} << 2 | c_{K+2} << 4.  One thread per K-mer walks the depth-3 tree of extensions; a 64-byte
// rank16 chunk carries the entries of all four bases of a block, so every tree node costs one or two loads.
__device__ __forceinline__ void lf4(const ulonglong2* r16, const uint64_t* fchr, uint64_t top, uint64_t bot, uint64_t t[4], uint64_t b[4]) {
	const ulonglong2* pt = r16 + (top >> 6) * 4; const ulonglong2* pb = r16 + (bot >> 6) * 4;
	const uint64_t mt = (1ull << (top & 63)) - 1ull, mb = (1ull << (bot & 63)) - 1ull;
	#pragma unroll
	for(int c = 0; c < 4; c++) {
		const ulonglong2 et = __ldg(pt + c); const ulonglong2 eb = (pb == pt) ? et : __ldg(pb + c);
		t[c] = fchr[c] + (et.x & 0x7fffffffffffffffull) + (uint64_t)__popcll(et.y & mt);
		b[c] = fchr[c] + (eb.x & 0x7fffffffffffffffull) + (uint64_t)__popcll(eb.y & mb);
	}
}
__global__ void __launch_bounds__(128) k_build_ftabd(IndexView v, int K, uint64_t n, uint8_t* out) {
	const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(f >= n) return;
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(v.rank16);
	const uint64_t* base = (K > v.ftab_chars) ? v.ftabk : v.ftab2;
	const uint64_t top = base[f * 2], bot = base[f * 2 + 1];
	uint64_t w0 = 0, w1 = 0;                                  // 64 x 2 bits, extension e in bits 2e
	if(bot > top) {
		uint64_t t1[4], b1[4]; lf4(r16, v.fchr, top, bot, t1, b1);
		for(int c0 = 0; c0 < 4; c0++) {
			if(b1[c0] <= t1[c0]) { for(int r = 0; r < 16; r++) { const int e = c0 | (r << 2); if(e < 32) w0 |= 1ull << (2 * e); else w1 |= 1ull << (2 * (e - 32)); } continue; }
			uint64_t t2[4], b2[4]; lf4(r16, v.fchr, t1[c0], b1[c0], t2, b2);
			for(int c1 = 0; c1 < 4; c1++) {
				if(b2[c1] <= t2[c1]) { for(int c2 = 0; c2 < 4; c2++) { const int e = c0 | (c1 << 2) | (c2 << 4); if(e < 32) w0 |= 2ull << (2 * e); else w1 |= 2ull << (2 * (e - 32)); } continue; }
				uint64_t t3[4], b3[4]; lf4(r16, v.fchr, t2[c1], b2[c1], t3, b3);
				for(int c2 = 0; c2 < 4; c2++) if(b3[c2] <= t3[c2]) { const int e = c0 | (c1 << 2) | (c2 << 4); if(e < 32) w0 |= 3ull << (2 * e); else w1 |= 3ull << (2 * (e - 32)); }
			}
		}
	}
	reinterpret_cast<ulonglong2*>(out)[f] = make_ulonglong2(w0, w1);
}

struct Blk { uint64_t m[4]; uint32_t pc[4]; };    // match masks of base c and their popcounts

__device__ __forceinline__ void blk_prep(Blk& q, const ulonglong2& a, const ulonglong2& b, uint64_t rep) {
	const uint64_t w[4] = {a.x, a.y, b.x, b.y};
	#pragma unroll
	for(int j = 0; j < 4; j++) {
		const uint64_t x = ~(w[j] ^ rep);
		q.m[j] = x & (x >> 1) & 0x5555555555555555ull;
		q.pc[j] = (uint32_t)__popcll(q.m[j]);
	}
}
// number of matches among the first off (0..127) rows of the block
__device__ __forceinline__ uint32_t blk_rank(const Blk& q, uint32_t off) {
	const uint32_t k = off >> 5, part = off & 31;
	uint32_t r = (k > 0 ? q.pc[0] : 0u) + (k > 1 ? q.pc[1] : 0u) + (k > 2 ? q.pc[2] : 0u);
	const uint64_t mk = k == 0 ? q.m[0] : (k == 1 ? q.m[1] : (k == 2 ? q.m[2] : q.m[3]));
	return r + (uint32_t)__popcll(shl64(mk, 64 - 2 * part));
}

// ---------------------------------------------------------------------------------------
// k_search_t: one thread per walk.  No cross-lane traffic: each lane reads the 32-byte rank sectors of
// its own top and bot rows.  Range and single-row steps share one code path, so the 32 independent
// walks of a warp diverge only on the restart / ftab paths, and those paths contain no blocking loads:
// the packed strand of the current task lives in registers (RW words, reads up to 32*RW bases), so the
// per-step base, the N test and the 10-mer ftab index are register extracts.  One loop iteration =
// one DRAM round trip for every walk of the warp.
// ---------------------------------------------------------------------------------------
template <int RW> struct ReadRegs {
	uint64_t rw[RW]; uint32_t nw[RW];
	__device__ __forceinline__ void load(const uint64_t* pk, const uint32_t* nm, uint32_t W) {
		#pragma unroll
		for(int k = 0; k < RW; k++) { rw[k] = (uint32_t)k < W ? __ldg(pk + k) : 0ull; nw[k] = (uint32_t)k < W ? __ldg(nm + k) : 0u; }
	}
	__device__ __forceinline__ uint64_t word(uint32_t k) const {
		uint64_t v = 0;
		#pragma unroll
		for(int q = 0; q < RW; q++) if((uint32_t)q == k) v = rw[q];
		return v;
	}
	__device__ __forceinline__ uint32_t nword(uint32_t k) const {
		uint32_t v = 0;
		#pragma unroll
		for(int q = 0; q < RW; q++) if((uint32_t)q == k) v = nw[q];
		return v;
	}
	// base at search depth p (4 = N)
	__device__ __forceinline__ int base(uint32_t p) const {
		const uint32_t k = p >> 5, sh = p & 31;
		return ((nword(k) >> sh) & 1u) ? 4 : (int)((word(k) >> (2 * sh)) & 3);
	}
	// bases p .. p+31 (base p in the low bits) and their N bits
	__device__ __forceinline__ void window(uint32_t p, uint64_t& win, uint32_t& nwin) const {
		const uint32_t k = p >> 5, sh = p & 31;
		win = shr64(word(k), 2 * sh) | shl64(word(k + 1), 64 - 2 * sh);
		nwin = (uint32_t)(((uint64_t)nword(k) | ((uint64_t)nword(k + 1) << 32)) >> sh);
	}
};

// COUNT: 0 = product; 1 = the reference's operation counters (jump tables off, so that the operation sequence is the
// reference's); 2 = the product's own load requests with every table live (what the roofline of *this* kernel is made of)
template <int COUNT, int RW>
__global__ void __launch_bounds__(kSearchThreads) k_search_t(const SearchArgs a) {      // 9 CTAs per SM at 52 registers; forcing 10 (48 registers) measured no faster: the DRAM gather rate is the limit
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(a.v.rank16);
	const ulonglong2* ftab2 = reinterpret_cast<const ulonglong2*>(a.v.ftab2);
	const ulonglong2* ftabk = reinterpret_cast<const ulonglong2*>(a.v.ftabk);
	const uint32_t fk = (COUNT == 1 || !a.v.ftabk) ? 0u : (uint32_t)a.v.ftabk_chars;   // counters follow the reference's op sequence
	const uint32_t fc = (uint32_t)a.v.ftab_chars;
	const unsigned long long* w8 = COUNT == 1 ? nullptr : reinterpret_cast<const unsigned long long*>(a.v.walk8);
	// death-depth table: only while every hit it can end (at most fd - 1 bases) stays below min_hitlen, i.e. is never resolved
	const uint32_t fd = (COUNT == 1 || !a.v.ftabd || a.p.min_hitlen < (uint32_t)a.v.ftabd_chars) ? 0u : (uint32_t)a.v.ftabd_chars;
	const uint32_t fdk = (uint32_t)a.v.ftabd_base;
	ReadRegs<RW> rd;
	uint64_t top = 0, bot = 0, fi = 0;
	uint32_t rlen = 0, tid = 0, cur = 0, dep = 0, offset = 0, nh = 0, nt = 0, slow_until = 0, fail_w = 0, fail_at = 0xffffffffu;
	bool nolong = true;      // no hit of this strand reaches min_hitlen (kListNoLong tells the per-unit kernels)
	int mode = M_NEED;
	unsigned long long c_ps = 0, c_ft = 0, c_sides = 0, c_lf = 0;
	unsigned long long q_r16 = 0, q_f2 = 0, q_fk = 0, q_w8 = 0, q_fd = 0;
	WarpPool pool; pool.base = pool.end = 0;
	bool more = true;      // warp-uniform: the global task counter is not exhausted yet

	auto emit = [&](uint64_t t, uint64_t b, uint32_t off, uint32_t len) {
		if(a.keep_short || len >= kLongLen) {
			if(nh < a.cap) { HitRec* h = a.hits + (size_t)tid * a.cap + nh; h->top = t; h->bot = b; h->bwoff = off; h->len = len; }
			else atomicExch(a.overflow, 1u);
			nh++;
		}
		nt++;
		if(len >= a.p.min_hitlen) nolong = false;
	};
	// searchForwardAndReverse restart policy (classifier.h:686-766): true = another partial search starts at cur
	auto after_hit = [&](uint32_t hlen) -> bool {
		bool done = cur >= rlen;
		if(!done) { if(hlen > a.p.increment) cur += 1; if(cur + a.p.min_hitlen >= rlen) done = true; }
		if(done) { a.nhits[tid] = nh_pack(nh, nt, nolong); mode = M_NEED; return false; }
		return true;
	};
	// partialSearch prologue (hi_aligner.h:939-982) at `cur`: ends in M_FTAB (fi set) or M_NEED
	auto start_search = [&]() {
		for(;;) {
			if(COUNT == 1) c_ps++;
			offset = cur;
			if(rlen - cur < fc) { emit(kOff, kOff, offset, rlen - offset); a.nhits[tid] = nh_pack(nh, nt, nolong); mode = M_NEED; return; }
			uint64_t win; uint32_t nwin; rd.window(cur, win, nwin);
			const uint32_t nbits = nwin & ((1u << fc) - 1u);
			if(nbits) {
				const uint32_t hl = (uint32_t)__ffs(nbits);
				cur += hl;
				emit(kOff, kOff, offset, hl);
				if(after_hit(hl)) continue; else return;
			}
			if(fd && rlen - cur >= fd && !(nwin & ((1u << fd) - 1u))) { fi = win & ((1ull << (2 * fd)) - 1ull); mode = M_FTABD; return; }
			if(fk && rlen - cur >= fk && !(nwin & ((1u << fk) - 1u))) { fi = win & ((1ull << (2 * fk)) - 1ull); mode = M_FTABK; return; }
			fi = win & ((1ull << (2 * fc)) - 1ull);
			mode = M_FTAB;
			return;
		}
	};
	auto hit_and_restart = [&]() {
		const uint32_t hl = dep - offset;
		emit(top, bot, offset, hl);
		cur = dep;
		if(after_hit(hl)) start_search();
	};

	for(;;) {
		// ---------------- hand out tasks (convergent point) ----------------
		{
			const bool want = mode == M_NEED;
			if(more) {
				unsigned long long t = 0;
				const bool got = pool_take(pool, want, a.task_ctr64, (unsigned long long)a.ntasks, a.chunk, t);
				if(want) {
					if(got) {
						tid = (uint32_t)t;
						const uint32_t per = 2u * (uint32_t)a.b.n_mates;
						const uint32_t unit = tid / per, rem = tid - unit * per;
						const int mate = (int)(rem >> 1);
						const uint8_t fl = a.b.flags ? a.b.flags[unit] : 3;
						nh = 0; nt = 0; rlen = a.b.len[mate][unit];
						if(!((fl >> mate) & 1) || rlen == 0) a.nhits[tid] = 0;          // filtered mate: stays M_NEED
						else { rd.load(a.pk + (size_t)tid * a.W, a.nm + (size_t)tid * a.W, a.W); cur = 0; slow_until = 0; fail_w = 0; fail_at = 0xffffffffu; nolong = true; start_search(); }
					} else mode = M_DONE;
				}
				if(__any_sync(0xffffffffu, want && !got)) more = false;      // global counter ran past the end
			} else if(want) mode = M_DONE;
		}
		if(!__any_sync(0xffffffffu, mode != M_DONE)) break;
		// ---------------- single fetch point ----------------
		// Every lane picks the address of the one 16-byte (aligned) piece it needs -- whatever table its walk is at -- plus
		// a second rank16 entry when a range straddles two 64-row blocks; then ONE predicated load instruction serves all
		// lanes (and one more the straddlers).  The loads are volatile asm so that the compiler cannot sink each table's
		// load into the branch that consumes it: it did, and a warp then paid one memory latency per table in play
		// instead of one per iteration (ncu, profiles/r02: stalls at the K-mer-table consumer with 4 of 32 lanes active).
		int c = 4;
		const bool lf = mode == M_LF;
		bool range = false, jump = false, known_fail = false;
		const void* p0 = nullptr; const void* p1 = nullptr; uint32_t sub = 0;
		if(mode == M_FTABD) {                                                // 2 bits of the death-depth table
			const uint64_t idx = ((fi & ((1ull << (2 * fdk)) - 1ull)) << 6) | (fi >> (2 * fdk));
			const uint64_t byte = idx >> 2;
			p0 = a.v.ftabd + (byte & ~15ull); sub = (uint32_t)(byte & 15) * 8u + (uint32_t)(idx & 3) * 2u;
			if(COUNT == 2) q_fd++;
		}
		else if(mode == M_FTAB) { p0 = ftab2 + fi; if(COUNT == 2) q_f2++; }                           // (top, bot) of the 10-mer
		else if(mode == M_FTABK) { p0 = ftabk + fi; if(COUNT == 2) q_fk++; }                          // (top, bot) of the K-mer
		else if(lf) {
			c = rd.base(dep);
			if(c <= 3) {
				range = (bot - top) != 1;
				const uint64_t width = bot - top;
				if(dep == fail_at && !range) known_fail = true;      // the walk8 entry already said that this step of the single row fails: no request
				else if(w8 && width <= (uint64_t)a.jump_w && (dep >= slow_until || width < (uint64_t)fail_w) && rlen - dep >= 8 && bot <= a.v.walk8_rows) {
					// Eight steps in one gather: while the walk holds a single row -- or a narrow range -- and the read's next eight
					// bases are the ones stored for its first AND its last row.  LF keeps the order of rows that continue with the
					// same base, so the range survives the eight steps intact exactly when both end rows do and their images are
					// still width - 1 apart (every row that drops out in between shortens that distance by one, nothing widens it).
					jump = true; p0 = w8 + (top & ~1ull); sub = (uint32_t)(top & 1); if(COUNT == 2) q_w8++;
					if(range) { const uint64_t last = bot - 1; sub |= (uint32_t)(last & 1) << 1; if((last & ~1ull) != (top & ~1ull)) { p1 = w8 + (last & ~1ull); if(COUNT == 2) q_w8++; } }
				} else {
					p0 = r16 + (top >> 6) * 4 + c;                          // (occ, bits): one request per rank query
					if(COUNT == 2) q_r16++;
					if(range && (bot >> 6) != (top >> 6)) { p1 = r16 + (bot >> 6) * 4 + c; if(COUNT == 2) q_r16++; }
				}
			}
		}
		ulonglong2 e = make_ulonglong2(0, 0), bq;
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u64 p, %2, 0;\n\t@p ld.global.nc.v2.u64 {%0, %1}, [%2];\n\t}" : "+l"(e.x), "+l"(e.y) : "l"(p0));
		bq = e;
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u64 p, %2, 0;\n\t@p ld.global.nc.v2.u64 {%0, %1}, [%2];\n\t}" : "+l"(bq.x), "+l"(bq.y) : "l"(p1));
		const ulonglong2 tq = e;
		// ---------------- consume ----------------
		if(mode == M_FTABD) {
			const uint32_t dv = (uint32_t)(((sub & 64u) ? e.y : e.x) >> (sub & 63u)) & 3u;
			if(dv == 0) {                                 // the (K+3)-mer occurs (or the K-mer does not): through the jump table
				fi &= (1ull << (2 * fdk)) - 1ull;
				mode = (fk && fdk == fk) ? M_FTABK : M_FTAB;
			} else {                                      // the range dies after K + v - 1 bases: that is the hit partialSearch reports
				const uint32_t hl = fdk + dv - 1u;
				emit(kUnk, kUnk, offset, hl);
				cur += hl;
				if(after_hit(hl)) start_search();
			}
		} else if(mode == M_FTABK) {
			if(e.y > e.x) {                               // same state partialSearch reaches after K bases
				top = e.x; bot = e.y; dep = cur + fk;
				if(dep < rlen) mode = M_LF; else hit_and_restart();
			} else { fi &= (1ull << (2 * fc)) - 1ull; mode = M_FTAB; }   // died between base fc and K: replay from the 10-mer
		} else if(mode == M_FTAB) {
			if(COUNT == 1) c_ft++;
			top = e.x; bot = e.y;
			dep = cur + fc;
			if(bot <= top) {                              // hi_aligner.h:971-982
				const uint32_t hl = dep - offset;
				emit(kOff, kOff, offset, hl);
				cur = dep;
				if(after_hit(hl)) start_search();
			} else if(dep < rlen) mode = M_LF;
			else hit_and_restart();
		} else if(jump) {
			uint64_t win; uint32_t nwin; rd.window(dep, win, nwin);
			const uint64_t w = (sub & 1u) ? e.y : e.x;                       // entry of the first row
			const uint64_t wl = (sub & 2u) ? bq.y : bq.x;                    // entry of the last row (the same entry for a single row)
			const uint64_t width = bot - top;
			if((w >> 56) == 8 && !(nwin & 0xffu) && !(((w >> 40) ^ win) & 0xffffull)
			   && (width == 1 || ((wl >> 40) == (w >> 40) && (wl & kWalkRowMask) - (w & kWalkRowMask) == width - 1))) {
				top = w & kWalkRowMask; bot = top + width; dep += 8;
				if(dep >= rlen) hit_and_restart();
			} else {      // some row leaves within the next eight steps: take them one by one (a narrower range may try again)
				slow_until = dep + 8; fail_w = (uint32_t)width;
				if(width == 1) {      // a single row: the entry tells which step ends the hit (first stored base that differs, an N, or the '$' row)
					const uint32_t diff = (uint32_t)(((w >> 40) ^ win) & 0xffffull), nb = nwin & 0xffu, nv = (uint32_t)(w >> 56) & 0xffu;
					uint32_t good = diff ? (uint32_t)(__ffs(diff) - 1) >> 1 : 8u;
					if(nb) good = min(good, (uint32_t)(__ffs(nb) - 1));
					good = min(good, nv);
					fail_at = good < 8u ? dep + good : 0xffffffffu;
				}
			}
		} else if(lf) {
			bool fail = c > 3 || known_fail;
			uint64_t t = 0, b = 0;
			if(!fail) {
				const uint32_t oT = (uint32_t)(top & 63), oB = (uint32_t)(bot & 63);
				t = a.v.fchr[c] + (tq.x & kOccMask) + (uint64_t)__popcll(tq.y & ((1ull << oT) - 1ull));
				if(range) b = a.v.fchr[c] + (bq.x & kOccMask) + (uint64_t)__popcll(bq.y & ((1ull << oB) - 1ull));
				else {                                    // mapLF1 bt2_idx.h:2910-2933: BWT[top] must be c ('$' has no bit)
					if(!((tq.y >> oT) & 1ull)) fail = true;
					b = t + 1;
				}
				if(b <= t) fail = true;
				if(COUNT == 1) {   // counters keep the reference's side geometry (384 rows per 128-byte side)
					uint64_t sT; uint32_t offT; row_locus(top, sT, offT);
					const bool same_side = !range || (bot - top) < (uint64_t)(384 - offT);
					c_lf += range ? 2 : 1; c_sides += same_side ? 1 : 2;
				}
			}
			if(fail) hit_and_restart();
			else { top = t; bot = b; dep++; if(dep >= rlen) hit_and_restart(); }
		}
	}
	if(COUNT == 1 && a.ctr) {
		atomicAdd(&a.ctr->partial_searches, c_ps); atomicAdd(&a.ctr->ftab_probes, c_ft);
		atomicAdd(&a.ctr->sides_search, c_sides); atomicAdd(&a.ctr->lf_steps, c_lf);
	}
	if(COUNT == 2 && a.ctr) {
		atomicAdd(&a.ctr->req_rank16, q_r16); atomicAdd(&a.ctr->req_ftab2, q_f2); atomicAdd(&a.ctr->req_ftabk, q_fk); atomicAdd(&a.ctr->req_walk8, q_w8); atomicAdd(&a.ctr->req_ftabd, q_fd);
	}
}

typedef void (*SearchKernel)(const SearchArgs);
// g: lanes per walk (2,4,8 = cooperative kernels on the sides; 16 = their G=1 instantiation);
// 1 / 100 = thread-per-walk kernels with the read in 4 / 10 register words
static SearchKernel search_kernel(int g, int count) {
	switch(g) {
		case 1: return count == 2 ? k_search_t<2, 4> : (count ? k_search_t<1, 4> : k_search_t<0, 4>);      // reads up to 128 bases
		case 101: return count == 2 ? k_search_t<2, 5> : (count ? k_search_t<1, 5> : k_search_t<0, 5>);    // reads up to 160 bases (2 x 150 bp runs)
		case 100: return count == 2 ? k_search_t<2, 10> : (count ? k_search_t<1, 10> : k_search_t<0, 10>);  // reads up to 320 bases
		case 16: return count ? k_search<true, 1> : k_search<false, 1>;   // generic template at G = 1 (A/B only)
		case 2: return count ? k_search<true, 2> : k_search<false, 2>;
		case 4: return count ? k_search<true, 4> : k_search<false, 4>;
		default: return count ? k_search<true, 8> : k_search<false, 8>;
	}
}

// =======================================================================================
// k_prep / k_rows / k_score (thread per unit)
// =======================================================================================
struct UnitArgs {
	IndexView v; Params p; BatchView b;
	HitRec* hits; uint32_t* nhits; uint32_t cap;
	uint32_t* nrows;            // per unit
	uint64_t* row_off;          // per unit: where its rows (ids, hit-map scratch, sparse records) start; handed out by k_prep
	unsigned long long* row_total;   // allocation counter = total rows of the batch
	uint64_t* rows; uint32_t* ids; uint64_t rows_cap;
	Entry* entries; TaxCnt* tcs; OutRec* recs_sparse; uint32_t* nout;
	unsigned int* overflow;
	Counters* ctr;
	HitRec* regen; uint32_t* regen_n; unsigned long long* regen_ctr; uint64_t regen_slots; uint32_t full_cap; uint32_t keep_short;
};

// One strand's whole greedy search by a single thread, with the device tables: the hits search_strand_scalar (cf_logic.h, the
// twin the CPU tests pin against the oracle) would produce, but reached the way k_search_t reaches them -- K-mer jump, one
// rank16 entry per step when top and bot share a block, eight bases per walk8 gather on a single row -- so that regenerating
// a list costs ~30 dependent gathers instead of ~250.  Used by k_prep only (lists whose short hits matter).
__device__ uint32_t search_strand_dev(const IndexView& v, const Params& p, const uint8_t* fw, uint32_t len, int strand, HitRec* hits, uint32_t cap) {
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(v.rank16);
	const ulonglong2* ftab2 = reinterpret_cast<const ulonglong2*>(v.ftab2);
	const ulonglong2* ftabk = reinterpret_cast<const ulonglong2*>(v.ftabk);
	const uint32_t fc = (uint32_t)v.ftab_chars, fk = v.ftabk ? (uint32_t)v.ftabk_chars : 0u;
	auto base = [&](uint32_t d) -> int { return seq_at(fw, len, strand, len - 1 - d); };     // the base consumed at search depth d
	uint32_t cur = 0, n = 0;
	if(len == 0) return 0;
	for(;;) {
		HitRec h; h.bwoff = cur; uint32_t new_cur; bool done = false;
		const uint32_t offset = cur;
		if(len - cur < fc) { h.top = h.bot = kOff; h.len = len - offset; new_cur = len; done = true; }
		else {
			uint32_t firstn = 0xffffffffu; uint64_t fi = 0;
			const uint32_t span = (fk && len - cur >= fk) ? fk : fc;
			for(uint32_t i = 0; i < span; i++) { const int c = base(cur + i); if(c > 3) { firstn = i; break; } fi |= (uint64_t)c << (2 * i); }
			if(firstn < fc) { new_cur = cur + firstn + 1; h.top = h.bot = kOff; h.len = new_cur - offset; done = new_cur >= len; }
			else {
				uint64_t top = 0, bot = 0; uint32_t dep = 0; bool have = false;
				if(span == fk && fk > fc && firstn == 0xffffffffu) { const ulonglong2 e = __ldg(ftabk + fi); if(e.y > e.x) { top = e.x; bot = e.y; dep = cur + fk; have = true; } }
				if(!have) { const ulonglong2 e = __ldg(ftab2 + (fi & ((1ull << (2 * fc)) - 1ull))); top = e.x; bot = e.y; dep = cur + fc; }
				if(bot <= top) { h.top = h.bot = kOff; h.len = dep - offset; new_cur = dep; done = dep >= len; }
				else {
					while(dep < len) {
						const int c = base(dep);
						if(c > 3) break;
						if(bot - top == 1) {
							if(v.walk8 && top < v.walk8_rows && len - dep >= 8) {        // eight bases in one gather
								const uint64_t e = __ldg(v.walk8 + top);
								bool ok = (e >> 56) == 8;
								for(uint32_t j = 0; ok && j < 8; j++) ok = base(dep + j) == (int)((e >> (40 + 2 * j)) & 3);
								if(ok) { top = e & kWalkRowMask; bot = top + 1; dep += 8; continue; }
							}
							const ulonglong2 e = __ldg(r16 + (top >> 6) * 4 + c);
							if(!((e.y >> (top & 63)) & 1ull)) break;                      // mapLF1: BWT[top] must be c ('$' has no bit)
							top = v.fchr[c] + (e.x & kOccMask) + (uint64_t)__popcll(e.y & ((1ull << (top & 63)) - 1ull)); bot = top + 1; dep++;
						} else {
							const ulonglong2 et = __ldg(r16 + (top >> 6) * 4 + c);
							const ulonglong2 eb = (bot >> 6) == (top >> 6) ? et : __ldg(r16 + (bot >> 6) * 4 + c);
							const uint64_t t = v.fchr[c] + (et.x & kOccMask) + (uint64_t)__popcll(et.y & ((1ull << (top & 63)) - 1ull));
							const uint64_t b = v.fchr[c] + (eb.x & kOccMask) + (uint64_t)__popcll(eb.y & ((1ull << (bot & 63)) - 1ull));
							if(b <= t) break;
							top = t; bot = b; dep++;
						}
					}
					h.top = top; h.bot = bot; h.len = dep - offset; new_cur = dep; done = dep >= len;
				}
			}
		}
		if(n < cap) hits[n] = h;
		n++;
		cur = new_cur;
		if(done) break;
		if(h.len > p.increment) cur += 1;
		if(cur + p.min_hitlen >= len) break;
	}
	return n;
}

// found[r][st] receives the number of hits the search found for the list (0 for a regenerated list); tpos[r] the index of
// the mate's first nhits word
__device__ __forceinline__ bool load_unit(const UnitArgs& a, uint32_t unit, UnitHits& u, const uint8_t* fw[2], uint32_t found[2][2], size_t tpos[2]) {
	const uint8_t fl = a.b.flags ? a.b.flags[unit] : 3;
	u.n_mates = 0;
	for(int m = 0; m < a.b.n_mates; m++) {
		if(!((fl >> m) & 1)) continue;
		const uint32_t len = a.b.len[m][unit];
		if(len == 0) continue;
		const int r = u.n_mates++;
		const size_t t0 = ((size_t)unit * a.b.n_mates + m) * 2;
		u.rdlen[r] = len; fw[r] = a.b.bases + a.b.off[m][unit]; tpos[r] = t0;
		const uint32_t raw[2] = {a.nhits[t0], a.nhits[t0 + 1]};
		for(int st = 0; st < 2; st++) {
			if(raw[st] & kListRegen) {                            // regenerated by an earlier k_prep pass over this batch
				const uint32_t slot = raw[st] & 0x3fffffffu;
				u.L[r][st] = a.regen + (size_t)slot * a.full_cap; u.n[r][st] = a.regen_n[slot]; found[r][st] = 0;
			} else {
				u.L[r][st] = a.hits + (t0 + st) * a.cap; u.n[r][st] = min(nh_stored(raw[st]), a.cap); found[r][st] = nh_found(raw[st]);
			}
		}
		// A strand list without a hit of min_hitlen bases matters only to the extension step, which needs such a hit on
		// BOTH strands (classifier.h:790-802); everything later (trimming within a list, strand choice, counting,
		// scoring) ignores or only shortens short hits.  So unless both strands have one, such a list is never read.
		if(!((raw[0] | raw[1]) & kListRegen) && ((raw[0] | raw[1]) & kListNoLong)) { if(raw[0] & kListNoLong) u.n[r][0] = 0; if(raw[1] & kListNoLong) u.n[r][1] = 0; }
	}
	return u.n_mates > 0;
}

// thread per unit: post-process + sort the hit lists, count the SA rows to resolve, take a slice of the row buffer
// (one atomic per warp) and write the rows with the scoring plan in their high bits.  EMIT_ONLY re-does only the last
// two steps on lists that are already final (after the row buffer had to grow).
template <int MINB, bool EMIT_ONLY>
__global__ void __launch_bounds__(128, MINB) k_prep(const UnitArgs a) {
	const uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = unit < a.b.n_units;
	UnitHits u; const uint8_t* fw[2]; uint32_t found[2][2]; size_t tpos[2];
	uint64_t rows = 0; bool have = false;
	if(live && (have = load_unit(a, unit, u, fw, found, tpos))) {
		if(EMIT_ONLY) { CountRows cr(a.p, u); for_each_visit(a.p, u, cr); rows = cr.rows; }
		else {
			Counters local; Counters* lc = nullptr;
			if(a.ctr) { memset(&local, 0, sizeof local); lc = &local; }
			// Only the long hits were stored (see kListRegen): where the short ones can matter, run the strand's search again
			// on this thread (search_strand_dev) into the side buffer, and let the list point there from now on.
			if(!a.keep_short) for(int r = 0; r < u.n_mates; r++) {
				const bool both = u.n[r][0] > 0 && u.n[r][1] > 0;
				for(int st = 0; st < 2; st++) {
					if(u.n[r][st] == 0 || !(both || found[r][st] > 16)) continue;
					const unsigned long long slot = atomicAdd(a.regen_ctr, 1ull);
					if(slot >= a.regen_slots) { atomicExch(a.overflow, 4u); continue; }      // the host grows the side buffer and re-runs the batch
					HitRec* L = a.regen + (size_t)slot * a.full_cap;
					const uint32_t n = search_strand_dev(a.v, a.p, fw[r], u.rdlen[r], st, L, a.full_cap);
					if(n > a.full_cap) atomicExch(a.overflow, 1u);
					u.L[r][st] = L; u.n[r][st] = min(n, a.full_cap);
					a.regen_n[slot] = u.n[r][st]; a.nhits[tpos[r] + st] = kListRegen | (uint32_t)slot;
				}
			}
			// Hits the death-depth table ended carry no SA range.  A short hit's range can matter only through the twin
			// removal (both strands in play, classifier.h:850-870) or through libstdc++'s tie permutation in lists long
			// enough for introsort (> 16 hits): recompute the ranges there, exactly as partialSearch would.
			for(int r = 0; r < u.n_mates; r++) {
				const bool both = u.n[r][0] > 0 && u.n[r][1] > 0;
				for(int st = 0; st < 2; st++) {
					if(!both && u.n[r][st] <= 16) continue;
					for(uint32_t i = 0; i < u.n[r][st]; i++) {
						HitRec& h = u.L[r][st][i];
						if(h.top != kUnk) continue;
						HitRec t; uint32_t nc; bool dn;
						partial_search_scalar(a.v, fw[r], u.rdlen[r], st, h.bwoff, t, nc, dn, nullptr);
						h.top = t.top; h.bot = t.bot;
					}
				}
			}
			for(int r = 0; r < u.n_mates; r++) post_search(a.v, a.p, fw[r], u.rdlen[r], u.L[r][0], u.n[r][0], u.L[r][1], u.n[r][1], lc);
			SortAndCount sc(a.p, u);
			for_each_visit(a.p, u, sc);
			rows = sc.rows;
			if(lc) {
				atomicAdd(&a.ctr->units, 1ull);
				if(local.ext_searches) {
					atomicAdd(&a.ctr->ext_searches, local.ext_searches); atomicAdd(&a.ctr->partial_searches, local.partial_searches);
					atomicAdd(&a.ctr->ftab_probes, local.ftab_probes); atomicAdd(&a.ctr->sides_search, local.sides_search); atomicAdd(&a.ctr->lf_steps, local.lf_steps);
				}
			}
		}
		if(rows > 0xFFFFFFFFull) rows = 0xFFFFFFFFull;
	}
	const unsigned lane = threadIdx.x & 31;
	uint64_t incl = rows;
	for(int d = 1; d < 32; d <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d); if((int)lane >= d) incl += t; }
	const uint64_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if(lane == 31 && warp_total) base = atomicAdd(a.row_total, (unsigned long long)warp_total);
	base = __shfl_sync(0xffffffffu, base, 31);
	if(!live) return;
	const uint64_t off = base + incl - rows;
	a.nrows[unit] = (uint32_t)rows; a.row_off[unit] = off;
	if(rows && off + rows <= a.rows_cap) { EmitRows er(a.p, u, a.rows + off); for_each_visit(a.p, u, er); }   // else: the host grows the buffer and re-runs EMIT_ONLY
}

static const int kLocalMap = 4;          // 16 measured slower (0.82 vs 0.63 ms per 2 M reads, profiles/r02_ab.txt): the bigger local frame costs more than the global scratch of the few heavy units
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_score(const UnitArgs a) {
	const uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x;
	if(unit >= a.b.n_units) return;
	uint32_t no = 0;
	const uint64_t off = a.row_off[unit], n = a.nrows[unit];
	if(n > 0 && *a.row_total <= a.rows_cap) {
		const uint8_t fl = a.b.flags ? a.b.flags[unit] : 3;
		int mates = 0;
		for(int m = 0; m < a.b.n_mates; m++) if(((fl >> m) & 1) && (m ? a.b.len[1][unit] : a.b.len[0][unit]) != 0) mates++;
		// the hit map of a unit with a handful of rows lives in thread-local memory (interleaved across the warp, L1-resident)
		// instead of the per-unit slice of the global scratch, whose 72-byte entries of neighbouring threads never share a sector
		Entry lmap[kLocalMap]; TaxCnt ltc[kLocalMap];
		Entry* map = n <= (uint64_t)kLocalMap ? lmap : a.entries + off;
		TaxCnt* tc = n <= (uint64_t)kLocalMap ? ltc : a.tcs + off;
		const uint32_t nmap = score_plan(a.v, a.p, a.rows + off, a.ids + off, n, map);
		no = reduce_and_emit(a.v, a.p, mates == 2, map, nmap, tc, a.recs_sparse + off);
	}
	a.nout[unit] = no;
}

// =======================================================================================
// scan (u32 -> exclusive u64, n+1 outputs) : 3 small kernels, 1024 elements per block
// =======================================================================================
static const int kScanBlock = 256, kScanPer = 4;   // 1024 per block

__global__ void __launch_bounds__(kScanBlock) k_scan_sums(const uint32_t* in, uint64_t n, uint64_t* bsum) {
	__shared__ uint64_t sh[kScanBlock];
	const uint64_t base = (uint64_t)blockIdx.x * kScanBlock * kScanPer + (uint64_t)threadIdx.x * kScanPer;
	uint64_t s = 0;
	for(int i = 0; i < kScanPer; i++) if(base + i < n) s += in[base + i];
	sh[threadIdx.x] = s; __syncthreads();
	for(int d = kScanBlock / 2; d > 0; d >>= 1) { if((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d]; __syncthreads(); }
	if(threadIdx.x == 0) bsum[blockIdx.x] = sh[0];
}
__global__ void k_scan_top(uint64_t* bsum, uint64_t nb, uint64_t* total) {   // single thread block, serial over blocks by chunks
	__shared__ uint64_t carry;
	__shared__ uint64_t sh[1024];
	if(threadIdx.x == 0) carry = 0;
	__syncthreads();
	for(uint64_t base = 0; base < nb; base += 1024) {
		const uint64_t i = base + threadIdx.x;
		const uint64_t vv = i < nb ? bsum[i] : 0;
		sh[threadIdx.x] = vv; __syncthreads();
		for(int d = 1; d < 1024; d <<= 1) { uint64_t t = (int)threadIdx.x >= d ? sh[threadIdx.x - d] : 0; __syncthreads(); sh[threadIdx.x] += t; __syncthreads(); }
		if(i < nb) bsum[i] = carry + sh[threadIdx.x] - vv;
		__syncthreads();
		if(threadIdx.x == 1023) carry += sh[1023];
		__syncthreads();
	}
	if(threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_apply(const uint32_t* in, uint64_t n, const uint64_t* bsum, const uint64_t* total, uint64_t* out) {
	__shared__ uint64_t sh[kScanBlock];
	const uint64_t base = (uint64_t)blockIdx.x * kScanBlock * kScanPer + (uint64_t)threadIdx.x * kScanPer;
	uint32_t vals[kScanPer]; uint64_t s = 0;
	for(int i = 0; i < kScanPer; i++) { vals[i] = base + i < n ? in[base + i] : 0; s += vals[i]; }
	sh[threadIdx.x] = s; __syncthreads();
	for(int d = 1; d < kScanBlock; d <<= 1) { uint64_t t = (int)threadIdx.x >= d ? sh[threadIdx.x - d] : 0; __syncthreads(); sh[threadIdx.x] += t; __syncthreads(); }
	uint64_t run = bsum[blockIdx.x] + sh[threadIdx.x] - s;
	for(int i = 0; i < kScanPer; i++) { if(base + i < n) out[base + i] = run; run += vals[i]; }
	if(blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

// =======================================================================================
// k_resolve : group of 8 lanes per SA row
// =======================================================================================
enum { R_DONE = 0, R_WALK = 1, R_SAMPLE = 2, R_NEED = 3 };
struct ResolveArgs {
	IndexView v; const uint64_t* rows; uint32_t* ids; uint16_t* ids16; const uint64_t* total; uint64_t rows_cap;
	unsigned long long* task_ctr; uint32_t chunk; Counters* ctr;
};

template <bool COUNT>
__global__ void __launch_bounds__(kSearchThreads) k_resolve(const ResolveArgs a) {
	const unsigned lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24, gmask = 0xFFu << gbase;
	const uint4* sides4 = reinterpret_cast<const uint4*>(a.v.sides);
	uint64_t n = *a.total; if(n > a.rows_cap) return;          // the row buffer was too small: slices have holes, the host re-runs the batch
	const uint64_t lowmask = ((uint64_t)1 << a.v.off_rate) - 1;
	uint64_t tnext = 0, tend = 0, idx = 0, row = 0;
	int mode = R_DONE;
	unsigned long long c_walk = 0, c_rows = 0;

	// classify `row` without memory: returns next mode; writes result for '$'
	auto settle = [&](uint64_t r) -> int {
		if(r == a.v.zoff) { if(gl == 0) a.ids[idx] = 0; return R_DONE; }
		if((r & lowmask) == 0) return R_SAMPLE;
		return R_WALK;
	};
	auto next_row = [&]() -> bool {       // loops until a row needs memory; false when out of work
		for(;;) {
			if(tnext >= tend) {
				unsigned long long base = 0;
				if(gl == 0) base = atomicAdd(a.task_ctr, (unsigned long long)a.chunk);
				base = __shfl_sync(gmask, base, gbase);
				if(base >= n) { mode = R_DONE; return false; }
				tnext = base; tend = base + a.chunk < n ? base + a.chunk : n;
			}
			idx = tnext++; row = a.rows[idx] & kRowMask;
			if(COUNT) c_rows++;
			mode = settle(row);
			if(mode != R_DONE) return true;
		}
	};
	next_row();
	while(__any_sync(0xffffffffu, mode != R_DONE)) {
		// ---------------- single fetch point ----------------
		uint4 da = make_uint4(0, 0, 0, 0); uint32_t bits = 0, samp = 0;
		uint64_t s = 0; uint32_t off = 0; bool chk = false;
		if(mode == R_WALK) {
			s = row / 384; off = (uint32_t)(row - s * 384);
			da = __ldg(sides4 + s * 8 + gl);
			chk = a.v.n_boundaries && a.v.last_boundary > 0 && row <= a.v.last_boundary;
			if(chk) bits = __ldg(a.v.bbits + ((row >> a.v.bshift) >> 5));
		} else if(mode == R_SAMPLE) {
			if(gl == 0) samp = a.v.sample32 ? __ldg(a.v.sample32 + (row >> a.v.off_rate)) : (uint32_t)__ldg(a.v.sample16 + (row >> a.v.off_rate));
		}
		// ---------------- consume ----------------
		if(mode == R_SAMPLE) {
			if(gl == 0) a.ids[idx] = samp;
			next_row();
		} else if(mode == R_WALK) {
			bool found = false;
			if(chk && (bits >> ((row >> a.v.bshift) & 31)) & 1u) {   // rare: binary search of the boundary rows
				uint32_t lo = 0, hi = a.v.n_boundaries;
				while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(a.v.brow[mid] < row) lo = mid + 1; else hi = mid; }
				if(lo < a.v.n_boundaries && a.v.brow[lo] == row) {
					found = true;
					if(gl == 0) a.ids[idx] = a.v.sample32 ? a.v.bseq[lo] : (uint32_t)(uint16_t)a.v.bseq[lo];
				}
			}
			if(found) next_row();
			else {
				const int c = group_char(da, off, gmask, gbase);
				const uint32_t rep = (uint32_t)c * 0x55555555u;
				int nn = (int)off - 64 * (int)gl; nn = nn < 0 ? 0 : (nn > 64 ? 64 : nn);
				uint32_t cnt = gl < 6 ? lane_count(da, rep, nn) : 0u;
				cnt = group_sum(cnt, gmask);
				const uint64_t occ = group_occ(da, c, gmask, gbase);
				uint64_t r = cnt;
				if(c == 0 && s == a.v.zside && a.v.zoffc < off) r--;
				row = a.v.fchr[c] + occ + r;
				if(COUNT) c_walk++;
				mode = settle(row);
				if(mode == R_DONE) next_row();
			}
		}
	}
	if(COUNT && gl == 0 && a.ctr) { atomicAdd(&a.ctr->walk_steps, c_walk); atomicAdd(&a.ctr->rows_resolved, c_rows); }
}

// ---------------------------------------------------------------------------------------
// k_resolve_t: one thread per SA row over the re-blocked index (same walk as k_resolve).
// ---------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(kSearchThreads, 8) k_resolve_t(const ResolveArgs a) {
	const uint64_t* blocks = a.v.blocks;
	const uint64_t zblk = a.v.zoff >> 7; const uint32_t zoffb = (uint32_t)(a.v.zoff & 127);
	uint64_t n = *a.total; if(n > a.rows_cap) return;          // the row buffer was too small: slices have holes, the host re-runs the batch
	const uint64_t lowmask = ((uint64_t)1 << a.v.off_rate) - 1;
	uint64_t idx = 0, row = 0;
	int mode = R_NEED;
	unsigned long long c_walk = 0, c_rows = 0;
	WarpPool pool; pool.base = pool.end = 0;
	bool more = true;
	// classify `row` without memory: next mode; a '$' row resolves on the spot
	auto settle = [&](uint64_t r) -> int {
		if(r == a.v.zoff) { a.ids[idx] = 0; return R_NEED; }
		if((r & lowmask) == 0) return R_SAMPLE;
		return R_WALK;
	};
	auto next_row = [&]() { mode = R_NEED; };
	for(;;) {
		{
			const bool want = mode == R_NEED;
			if(more) {
				unsigned long long t = 0;
				const bool got = pool_take(pool, want, a.task_ctr, (unsigned long long)n, a.chunk, t);
				if(want) {
					if(got) { idx = t; row = a.rows[idx] & kRowMask; if(COUNT) c_rows++; mode = settle(row); }
					else mode = R_DONE;
				}
				if(__any_sync(0xffffffffu, want && !got)) more = false;
			} else if(want) mode = R_DONE;
		}
		if(!__any_sync(0xffffffffu, mode != R_DONE)) break;
		ulonglong2 o0, o1, d0, d1; o0 = o1 = d0 = d1 = make_ulonglong2(0, 0);
		uint32_t bits = 0, samp = 0; bool chk = false;
		const uint64_t blk = row >> 7; const uint32_t off = (uint32_t)(row & 127);
		if(mode == R_WALK) {
			const ulonglong2* p = reinterpret_cast<const ulonglong2*>(blocks + blk * 8);
			o0 = __ldg(p); o1 = __ldg(p + 1); d0 = __ldg(p + 2); d1 = __ldg(p + 3);
			chk = a.v.n_boundaries && a.v.last_boundary > 0 && row <= a.v.last_boundary;
			if(chk) bits = __ldg(a.v.bbits + ((row >> a.v.bshift) >> 5));
		} else if(mode == R_SAMPLE) {
			samp = a.v.sample32 ? __ldg(a.v.sample32 + (row >> a.v.off_rate)) : (uint32_t)__ldg(a.v.sample16 + (row >> a.v.off_rate));
		}
		if(mode == R_SAMPLE) { a.ids[idx] = samp; next_row(); }
		else if(mode == R_WALK) {
			bool found = false;
			if(chk && ((bits >> ((row >> a.v.bshift) & 31)) & 1u)) {
				uint32_t lo = 0, hi = a.v.n_boundaries;
				while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(a.v.brow[mid] < row) lo = mid + 1; else hi = mid; }
				if(lo < a.v.n_boundaries && a.v.brow[lo] == row) { found = true; a.ids[idx] = a.v.sample32 ? a.v.bseq[lo] : (uint32_t)(uint16_t)a.v.bseq[lo]; }
			}
			if(found) next_row();
			else {
				const uint32_t k = off >> 5;
				const uint64_t wk = k == 0 ? d0.x : (k == 1 ? d0.y : (k == 2 ? d1.x : d1.y));
				const int c = (int)((wk >> (2 * (off & 31))) & 3);
				const uint64_t occ = c == 0 ? o0.x : (c == 1 ? o0.y : (c == 2 ? o1.x : o1.y));
				Blk q; blk_prep(q, d0, d1, (uint64_t)c * 0x5555555555555555ull);
				uint64_t r = blk_rank(q, off);
				if(c == 0 && blk == zblk && zoffb < off) r--;
				row = a.v.fchr[c] + occ + r;
				if(COUNT) c_walk++;
				mode = settle(row);
			}
		}
	}
	if(COUNT && a.ctr) { atomicAdd(&a.ctr->walk_steps, c_walk); atomicAdd(&a.ctr->rows_resolved, c_rows); }
}

// ---------------------------------------------------------------------------------------
// k_resolve_c: 4 lanes per SA row over rank16.  Lane j fetches the entry of base j, so the 64-byte
// chunk of a block arrives with ONE load request per walk step; the lane whose indicator bit is set
// at the row knows BWT[row] and its own LF value is the next row.  The genome-boundary prefilter is the
// flag bit in A's occ word (no separate bitmap load).
// ---------------------------------------------------------------------------------------
// IDENT: rows are 0..n-1 themselves and results go to the (16- or 32-bit) resolve table -- used once at index load
template <bool COUNT, bool IDENT>
__global__ void __launch_bounds__(kSearchThreads) k_resolve_c(const ResolveArgs a) {
	const unsigned lane = threadIdx.x & 31, gl = lane & 3, gbase = lane & 28, gmask = 0xFu << gbase;
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(a.v.rank16);
	uint64_t n = *a.total; if(n > a.rows_cap) return;          // the row buffer was too small: slices have holes, the host re-runs the batch
	const uint64_t lowmask = ((uint64_t)1 << a.v.off_rate) - 1;
	uint64_t idx = 0, row = 0;
	int mode = R_NEED;
	unsigned long long c_walk = 0, c_rows = 0;
	WarpPool pool; pool.base = pool.end = 0;
	bool more = true;
	auto put = [&](uint32_t v) { if(IDENT && a.ids16) a.ids16[idx] = (uint16_t)v; else a.ids[idx] = v; };
	auto settle = [&](uint64_t r) -> int {
		if(r == a.v.zoff) { if(gl == 0) put(0); return R_NEED; }
		if((r & lowmask) == 0) return R_SAMPLE;
		return R_WALK;
	};
	for(;;) {
		{
			const bool want = mode == R_NEED;
			if(more) {
				unsigned long long t = 0;
				bool got = pool_take(pool, want && gl == 0, a.task_ctr, (unsigned long long)n, a.chunk, t);
				t = __shfl_sync(0xffffffffu, t, gbase); got = __shfl_sync(0xffffffffu, (int)got, gbase) != 0;
				if(want) {
					if(got) { idx = t; row = IDENT ? idx : (a.rows[idx] & kRowMask); if(COUNT && gl == 0) c_rows++; mode = settle(row); }
					else mode = R_DONE;
				}
				if(__any_sync(0xffffffffu, want && !got)) more = false;
			} else if(want) mode = R_DONE;
		}
		if(!__any_sync(0xffffffffu, mode != R_DONE)) break;
		ulonglong2 e = make_ulonglong2(0, 0); uint32_t samp = 0;
		if(mode == R_WALK) e = __ldg(r16 + (row >> 6) * 4 + gl);
		else if(mode == R_SAMPLE && gl == 0) samp = a.v.sample32 ? __ldg(a.v.sample32 + (row >> a.v.off_rate)) : (uint32_t)__ldg(a.v.sample16 + (row >> a.v.off_rate));
		if(mode == R_SAMPLE) { if(gl == 0) put(samp); mode = R_NEED; }
		else if(mode == R_WALK) {
			const uint32_t off = (uint32_t)(row & 63);
			const unsigned flagged = __shfl_sync(gmask, (unsigned)(e.x >> 63), gbase);     // A's entry carries the boundary flag
			bool found = false;
			if(flagged && a.v.last_boundary > 0 && row <= a.v.last_boundary) {
				uint32_t lo = 0, hi = a.v.n_boundaries;
				while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(a.v.brow[mid] < row) lo = mid + 1; else hi = mid; }
				if(lo < a.v.n_boundaries && a.v.brow[lo] == row) { found = true; if(gl == 0) put(a.v.sample32 ? a.v.bseq[lo] : (uint32_t)(uint16_t)a.v.bseq[lo]); }
			}
			if(found) mode = R_NEED;
			else {
				const uint64_t mine = a.v.fchr[gl] + (e.x & kOccMask) + (uint64_t)__popcll(e.y & ((1ull << off) - 1ull));
				const unsigned who = (__ballot_sync(gmask, (e.y >> off) & 1ull) >> gbase) & 0xFu;      // exactly one base owns the row
				const int src = __ffs(who) - 1;
				row = __shfl_sync(gmask, mine, gbase + (src < 0 ? 0 : src));
				if(COUNT && gl == 0) c_walk++;
				mode = settle(row);
			}
		}
	}
	if(COUNT && gl == 0 && a.ctr) { atomicAdd(&a.ctr->walk_steps, c_walk); atomicAdd(&a.ctr->rows_resolved, c_rows); }
}

// walk8: for every SA row r the state of eight successive mapLF1 steps (bt2_idx.h:2910): the bases BWT[r0..r7]
// (2 bits each, first step in the low bits) and the row reached, packed as row | bases << 40 | n_valid << 56.
// n_valid < 8 when the walk meets the '$' row.  While a search holds a single row and the next eight read
// bases equal the stored ones, eight dependent rank gathers collapse into this one 8-byte gather.
// 4 lanes per row (lane j fetches base j's rank16 entry: the 64-byte chunk is one request).
__global__ void __launch_bounds__(kSearchThreads) k_build_walk8(IndexView v, uint64_t nrows, uint64_t* out) {
	const unsigned lane = threadIdx.x & 31, gl = lane & 3, gbase = lane & 28, gmask = 0xFu << gbase;
	const ulonglong2* r16 = reinterpret_cast<const ulonglong2*>(v.rank16);
	const uint64_t ngroups = (uint64_t)gridDim.x * (kSearchThreads / 4);
	const uint64_t per = (nrows + ngroups - 1) / ngroups;
	const uint64_t g = (uint64_t)blockIdx.x * (kSearchThreads / 4) + (threadIdx.x >> 2);
	// consecutive rows per group keep the first gathers of neighbouring rows in the same 64-row chunk
	for(uint64_t idx = g * per; idx < min(nrows, (g + 1) * per); idx++) {
		uint64_t row = idx, chars = 0; uint32_t nv = 0;
		for(; nv < 8; nv++) {
			if(row == v.zoff) break;
			const ulonglong2 e = __ldg(r16 + (row >> 6) * 4 + gl);
			const uint32_t off = (uint32_t)(row & 63);
			const uint64_t mine = v.fchr[gl] + (e.x & kOccMask) + (uint64_t)__popcll(e.y & ((1ull << off) - 1ull));
			const unsigned who = (__ballot_sync(gmask, (e.y >> off) & 1ull) >> gbase) & 0xFu;
			const int src = __ffs(who) - 1;
			if(src < 0) break;
			chars |= (uint64_t)src << (2 * nv);
			row = __shfl_sync(gmask, mine, gbase + src);
		}
		if(gl == 0) out[idx] = (row & kWalkRowMask) | (chars << 40) | ((uint64_t)nv << 56);
	}
}

// Resolve by table: the sequence id of every SA row was precomputed at index load (k_resolve_c<.,true>), so
// resolving a row is one 2- or 4-byte gather instead of a ~8-step dependent walk.
__global__ void __launch_bounds__(256) k_lookup(const ResolveArgs a) {
	uint64_t n = *a.total; if(n > a.rows_cap) return;          // the row buffer was too small: slices have holes, the host re-runs the batch
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for(uint64_t k = i; k < n; k += stride) { const uint64_t r = a.rows[k] & kRowMask; a.ids[k] = a.v.rtab32 ? __ldg(a.v.rtab32 + r) : (uint32_t)__ldg(a.v.rtab16 + r); }
}

// =======================================================================================
// k_compact
// =======================================================================================
__global__ void __launch_bounds__(128) k_compact(uint32_t n_units, const uint64_t* row_off, const uint64_t* out_off,
                                                 const OutRec* sparse, OutRec* dense, uint32_t* rec_off32, uint64_t dense_cap, unsigned int* overflow) {
	const uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x;
	if(unit > n_units) return;
	if(unit == n_units) { rec_off32[unit] = (uint32_t)out_off[unit]; return; }
	const uint64_t o = out_off[unit], n = out_off[unit + 1] - o;
	rec_off32[unit] = (uint32_t)o;
	if(o + n > dense_cap) { if(n) atomicExch(overflow, 3u); return; }
	const uint64_t src = row_off[unit];
	for(uint64_t i = 0; i < n; i++) dense[o + i] = sparse[src + i];
}

// =======================================================================================
// test hook kernels
// =======================================================================================
__global__ void k_test_lf(IndexView v, const uint64_t* rows, const uint8_t* chars, uint64_t n, uint64_t* out) {
	const unsigned lane = threadIdx.x & 31, gl = lane & 7, gbase = lane & 24, gmask = 0xFFu << gbase;
	const uint4* sides4 = reinterpret_cast<const uint4*>(v.sides);
	const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
	const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
	const uint64_t iters = (n + ngroups - 1) / ngroups;          // uniform trip count: all groups of a warp stay in lockstep
	for(uint64_t it = 0; it < iters; it++) {
		const uint64_t i = g + it * ngroups;
		const bool act = i < n;
		const uint64_t row = act ? rows[i] : 0;
		const uint64_t s = row / 384; const uint32_t off = (uint32_t)(row - s * 384);
		const uint4 da = __ldg(sides4 + s * 8 + gl);
		int c = act ? chars[i] : 0;
		const int rowc = group_char(da, off, gmask, gbase);
		if(c > 3) c = rowc;
		const uint32_t rep = (uint32_t)c * 0x55555555u;
		int nn = (int)off - 64 * (int)gl; nn = nn < 0 ? 0 : (nn > 64 ? 64 : nn);
		uint32_t cnt = gl < 6 ? lane_count(da, rep, nn) : 0u;
		cnt = group_sum(cnt, gmask);
		const uint64_t occ = group_occ(da, c, gmask, gbase);
		uint64_t r = cnt;
		if(c == 0 && s == v.zside && v.zoffc < off) r--;
		if(act && gl == 0) out[i] = v.fchr[c] + occ + r;
	}
}

// =======================================================================================
// host side: index replica, contexts, batches
// =======================================================================================
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
	va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code;
}
#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) return fail(CFB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

extern "C" const char* cfb_last_error(void) { return g_err; }
extern "C" const char* cfb_version(void) { return "cfb200 0.1 (sm_100a)"; }

template <class T> struct DBuf {     // growable device buffer
	T* p = nullptr; size_t cap = 0;
	cudaError_t ensure(size_t n) {
		if(n <= cap) return cudaSuccess;
		if(p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 16;
		cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
		if(e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};
template <class T> struct HBuf {     // growable pinned host buffer
	T* p = nullptr; size_t cap = 0;
	cudaError_t ensure(size_t n) {
		if(n <= cap) return cudaSuccess;
		if(p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 16;
		cudaError_t e = cudaMallocHost((void**)&p, want * sizeof(T));
		if(e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if(p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct cfb_index {
	HostIndex h;
	int device = -1;
	IndexView view;            // device pointers (taxonomy-independent part)
	std::vector<void*> dptrs;
	uint64_t device_bytes = 0;
	int sm_count = 0;
	cfb_index_tables tables;
	cfb_index() { memset(&tables, 0, sizeof tables); }
};

static int upload(cfb_index* ix, const void* src, size_t bytes, const void** dst) {
	void* d = nullptr;
	CK(cudaMalloc(&d, bytes ? bytes : 16));
	if(bytes) CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
	ix->dptrs.push_back(d); ix->device_bytes += bytes;
	*dst = d;
	return CFB_OK;
}

// file[off, off+bytes) -> device memory through a ring of pinned buffers: parallel preads fill one buffer while
// the previous one is on its way over PCIe.  No host copy of the array is kept.
static int stream_to_device(cfb_index* ix, const std::string& path, uint64_t off, uint64_t bytes, const void** dst) {
	void* d = nullptr;
	CK(cudaMalloc(&d, bytes ? bytes + 64 : 64));
	ix->dptrs.push_back(d); ix->device_bytes += bytes; *dst = d;
	if(bytes == 0) return CFB_OK;
	const int fd = open(path.c_str(), O_RDONLY);
	if(fd < 0) return fail(CFB_EIO, "could not open %s", path.c_str());
	const size_t kBuf = 64u << 20; const int kRing = 3, kThreads = 8;
	uint8_t* hb[kRing] = {nullptr, nullptr, nullptr}; cudaEvent_t ev[kRing]; cudaStream_t st = nullptr;
	int rc = CFB_OK;
	for(int i = 0; i < kRing; i++) { if(cudaMallocHost((void**)&hb[i], kBuf) != cudaSuccess || cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) rc = fail(CFB_ENOMEM, "pinned staging buffers"); }
	if(rc == CFB_OK && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) rc = fail(CFB_ECUDA, "stream");
	uint64_t done = 0; int k = 0;
	while(rc == CFB_OK && done < bytes) {
		const size_t n = (size_t)std::min<uint64_t>(kBuf, bytes - done);
		const int b = k % kRing;
		if(k >= kRing && cudaEventSynchronize(ev[b]) != cudaSuccess) { rc = fail(CFB_ECUDA, "event"); break; }
		std::atomic<bool> bad(false);
		auto piece = [&](int t) {
			const size_t lo = n * t / kThreads, hi = n * (t + 1) / kThreads; size_t got = lo;
			while(got < hi) { const ssize_t r = pread(fd, hb[b] + got, hi - got, (off_t)(off + done + got)); if(r <= 0) { bad = true; return; } got += (size_t)r; }
		};
		std::vector<std::thread> th;
		for(int t = 1; t < kThreads; t++) th.emplace_back(piece, t);
		piece(0);
		for(size_t t = 0; t < th.size(); t++) th[t].join();
		if(bad) { rc = fail(CFB_EIO, "short read in %s", path.c_str()); break; }
		if(cudaMemcpyAsync((uint8_t*)d + done, hb[b], n, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaEventRecord(ev[b], st) != cudaSuccess) { rc = fail(CFB_ECUDA, "H2D of %s", path.c_str()); break; }
		done += n; k++;
	}
	if(st) { if(cudaStreamSynchronize(st) != cudaSuccess && rc == CFB_OK) rc = fail(CFB_ECUDA, "H2D of %s", path.c_str()); cudaStreamDestroy(st); }
	for(int i = 0; i < kRing; i++) { if(hb[i]) cudaFreeHost(hb[i]); cudaEventDestroy(ev[i]); }
	close(fd);
	return rc;
}

// CK inside the loader: the half-built replica is released before the error is returned
#define CKX(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { const int rc_ = fail(CFB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); cfb_index_free(ix); return rc_; } } while(0)
extern "C" int cfb_index_load_ex(const char* basename, int device, uint32_t flags, cfb_index** out) {
	if(!basename || !out) return fail(CFB_EINVAL, "cfb_index_load: null argument");
	cfb_index* ix = new cfb_index();
	std::string err = load_cf_index(basename, ix->h, /*defer_bulk=*/device >= 0);
	if(!err.empty()) { delete ix; return fail(CFB_EIO, "%s", err.c_str()); }
	const HostIndex& h = ix->h;
	ix->device = device;
	if(device >= 0) {
		if(h.line_rate != 7) { const int lr = h.line_rate; delete ix; return fail(CFB_EFORMAT, "index lineRate %d unsupported: the sm_100a kernels require 128-byte sides (centrifuge-build default --linerate 7)", lr); }
		if(h.ftab_chars > 15) { const int fc = h.ftab_chars; delete ix; return fail(CFB_EFORMAT, "ftabChars %d unsupported", fc); }
		int ndev = 0;
		if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { delete ix; return fail(CFB_ENODEV, "no CUDA device %d (found %d); this library has no CPU fallback", device, ndev); }
		cudaDeviceProp prop;
		if(cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ix; return fail(CFB_ENODEV, "cannot use CUDA device %d", device); }
		ix->sm_count = prop.multiProcessorCount;
		IndexView& v = ix->view; memset(&v, 0, sizeof v);
		int rc;
		#define UP(field, vec, T) if((rc = upload(ix, (vec).data(), (vec).size() * sizeof(T), (const void**)&v.field)) != CFB_OK) { cfb_index_free(ix); return rc; }
		if((rc = stream_to_device(ix, std::string(basename) + ".1.cf", h.sides_file_off, h.num_sides * h.side_sz, (const void**)&v.sides)) != CFB_OK) { cfb_index_free(ix); return rc; }
		ix->tables.sides_bytes = h.num_sides * h.side_sz; ix->tables.sample_bytes = h.offs_len * (h.wide_sample ? 4 : 2);
		UP(ftab, h.ftab, uint64_t) UP(eftab, h.eftab, uint64_t)
		if((rc = stream_to_device(ix, std::string(basename) + ".2.cf", h.sample_file_off, h.offs_len * (h.wide_sample ? 4 : 2),
		                          h.wide_sample ? (const void**)&v.sample32 : (const void**)&v.sample16)) != CFB_OK) { cfb_index_free(ix); return rc; }
		UP(brow, h.brow, uint64_t) UP(bseq, h.bseq, uint32_t) UP(bbits, h.bbits, uint32_t)
		UP(seq_taxid, h.seq_taxid, uint64_t) UP(seq_path, h.seq_path, int32_t) UP(paths, h.paths, uint64_t)
		#undef UP
		v.len = h.len; v.zoff = h.zoff; v.zside = h.zoff / 384; v.zoffc = (uint32_t)(h.zoff % 384);
		for(int i = 0; i < 4; i++) v.fchr[i] = h.fchr[i];
		v.last_boundary = h.last_boundary; v.num_sides = h.num_sides;
		v.n_boundaries = (uint32_t)h.brow.size(); v.n_seqs = (uint32_t)h.seq_taxid.size();
		v.off_rate = h.off_rate; v.ftab_chars = h.ftab_chars; v.bshift = h.bshift;
		if(getenv("CFB_LEGACY_LAYOUTS")) {   // 64-byte blocks: only for the k_resolve_t A/B variant (CFB_RESOLVE=1)
			uint64_t* blk = nullptr; const uint64_t nb = h.num_sides * 3;
			CKX(cudaMalloc((void**)&blk, (nb + 1) * 64));
			ix->dptrs.push_back(blk); ix->device_bytes += (nb + 1) * 64;
			k_build_blocks<<<(unsigned)((nb + 1 + 255) / 256), 256>>>(v.sides, h.num_sides, v.zside, v.zoffc, blk);
			CKX(cudaDeviceSynchronize());
			v.blocks = blk; v.num_blocks = nb;
		}
		if(getenv("CFB_LEGACY_LAYOUTS")) {   // 32-byte rank sectors: superseded by rank16, kept for A/B only
			uint64_t* rv = nullptr; const uint64_t nb = h.num_sides * 2;
			CKX(cudaMalloc((void**)&rv, (nb + 1) * 128));
			ix->dptrs.push_back(rv); ix->device_bytes += (nb + 1) * 128;
			k_build_rankv<<<(unsigned)((nb + 1 + 255) / 256), 256>>>(v.sides, h.num_sides, v.zside, v.zoffc, rv);
			CKX(cudaDeviceSynchronize());
			v.rankv = rv;
		}
		{   // 16-byte rank entries + fused ftab for the walk kernels
			uint64_t* r16 = nullptr; const uint64_t nb = h.num_sides * 6;
			CKX(cudaMalloc((void**)&r16, (nb + 1) * 64));
			ix->dptrs.push_back(r16); ix->device_bytes += (nb + 1) * 64;
			k_build_rank16<<<(unsigned)((nb + 1 + 255) / 256), 256>>>(v.sides, h.num_sides, v.zside, v.zoffc, r16);
			if(v.n_boundaries) k_mark_boundaries<<<(v.n_boundaries + 255) / 256, 256>>>(v.brow, v.n_boundaries, r16);
			uint64_t* f2 = nullptr; const uint64_t nf = h.ftab_len - 1;
			CKX(cudaMalloc((void**)&f2, nf * 16));
			ix->dptrs.push_back(f2); ix->device_bytes += nf * 16;
			k_build_ftab2<<<(unsigned)((nf + 255) / 256), 256>>>(v, nf, f2);
			CKX(cudaDeviceSynchronize());
			v.rank16 = r16; v.ftab2 = f2;
			ix->tables.rank16_bytes = (nb + 1) * 64; ix->tables.ftab2_bytes = nf * 16;
			// The file's sides have served their purpose (rank16 holds the same information, and the rare scalar LF of the
			// extension step reads rank16 too): free them unless the sides-based A/B kernels and test hooks are wanted
			// (small indexes keep them; CFB_KEEP_SIDES=1 / CFB_LEGACY_LAYOUTS=1 force it).
			if(ix->tables.sides_bytes > (256ull << 20) && !getenv("CFB_KEEP_SIDES") && !getenv("CFB_LEGACY_LAYOUTS")) {
				cudaFree((void*)v.sides);
				for(size_t i = 0; i < ix->dptrs.size(); i++) if(ix->dptrs[i] == (void*)v.sides) { ix->dptrs.erase(ix->dptrs.begin() + i); break; }
				ix->device_bytes -= ix->tables.sides_bytes; ix->tables.sides_bytes = 0; v.sides = nullptr;
			}
			// HBM budget of the derived tables: what is free now minus the head-room the batch buffers need (24 GB by default:
			// 16 slots of 0.5 M reads at ~3.5 KB each; CFB_HBM_HEADROOM_GB).  Tables are built in the order of gathers saved per
			// byte -- K-mer jump table, resolve table, death-depth table, walk8 -- each only if it fits what is left; walk8, whose rows are hit
			// uniformly, may cover just a prefix of the rows (a jump needs an entry for the row it starts from only).
			size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
			double head_gb = 24.0; { const char* e = getenv("CFB_HBM_HEADROOM_GB"); if(e) head_gb = atof(e); }
			const uint64_t headroom = std::min<uint64_t>((uint64_t)(head_gb * 1073741824.0), free_b / 2);
			auto budget = [&]() -> uint64_t { size_t f = 0, t = 0; cudaMemGetInfo(&f, &t); return f > headroom ? f - headroom : 0; };
			// extended jump table: K = largest value with 4^K <= len/4 (most K-mers occur), capped at 15 and by the budget
			int K = 0;
			{ const char* e = getenv("CFB_FTABK"); if(e) K = atoi(e); else { K = h.ftab_chars; while(K < 15 && (4ull << (2 * K)) <= h.len / 4) K++; } }
			while(K > h.ftab_chars && (16ull << (2 * K)) > budget() / 2) K--;
			if(K > h.ftab_chars && K <= 16) {
				uint64_t* fk = nullptr; const uint64_t nk = 1ull << (2 * K);
				CKX(cudaMalloc((void**)&fk, nk * 16));
				ix->dptrs.push_back(fk); ix->device_bytes += nk * 16;
				k_build_ftabk<<<(unsigned)((nk + 255) / 256), 256>>>(v, K, nk, fk);
				CKX(cudaDeviceSynchronize());
				v.ftabk = fk; v.ftabk_chars = K;
				ix->tables.ftabk_bytes = nk * 16; ix->tables.ftabk_chars = K;
			}
			// resolve table: sequence id of every SA row (walked once here)
			{
				const char* e = (flags & CFB_LOAD_NO_RESOLVE_TABLE) ? "0" : getenv("CFB_RESOLVE_TABLE");
				const uint64_t nrows = h.len + 1, esz = h.wide_sample ? 4 : 2;
				if(!(e && e[0] == '0') && nrows * esz + 16 <= budget()) {
					void* tab = nullptr; unsigned long long* sc = nullptr;
					CKX(cudaMalloc(&tab, nrows * esz + 16)); CKX(cudaMalloc((void**)&sc, 16));
					ix->dptrs.push_back(tab); ix->device_bytes += nrows * esz;
					const unsigned long long init[2] = {0ull, (unsigned long long)nrows};
					CKX(cudaMemcpy(sc, init, 16, cudaMemcpyHostToDevice));
					ResolveArgs ra; ra.v = v; ra.rows = nullptr; ra.ids = h.wide_sample ? (uint32_t*)tab : nullptr; ra.ids16 = h.wide_sample ? nullptr : (uint16_t*)tab;
					ra.total = (const uint64_t*)(sc + 1); ra.rows_cap = nrows; ra.task_ctr = sc; ra.chunk = 256; ra.ctr = nullptr;
					int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resolve_c<false, true>, kSearchThreads, 0);
					k_resolve_c<false, true><<<prop.multiProcessorCount * std::max(occ, 1), kSearchThreads>>>(ra);
					CKX(cudaDeviceSynchronize());
					cudaFree(sc);
					if(h.wide_sample) v.rtab32 = (const uint32_t*)tab; else v.rtab16 = (const uint16_t*)tab;
					ix->tables.resolve_table_bytes = nrows * esz; ix->tables.resolve_entry_bytes = (int32_t)esz;
				}
			}
			// death-depth table over the (K+3)-mers of the jump table in use (the K-mer table, else the 10-mer table): same
			// size as a K-mer table of that K (16 bytes per K-mer)
			{
				const char* e = getenv("CFB_FTABD");
				const int Kb = v.ftabk ? v.ftabk_chars : h.ftab_chars;
				const uint64_t nk = 1ull << (2 * Kb);
				if(!(e && e[0] == '0') && Kb + 3 <= 20 && nk * 16 <= budget()) {
					uint8_t* fd = nullptr;
					CKX(cudaMalloc((void**)&fd, nk * 16));
					ix->dptrs.push_back(fd); ix->device_bytes += nk * 16;
					k_build_ftabd<<<(unsigned)((nk + 127) / 128), 128>>>(v, Kb, nk, fd);
					CKX(cudaDeviceSynchronize());
					v.ftabd = fd; v.ftabd_chars = Kb + 3; v.ftabd_base = Kb;
					ix->tables.ftabd_bytes = nk * 16; ix->tables.ftabd_chars = Kb + 3;
				}
			}
			// walk8: eight single-row LF steps per gather, for as many rows as the budget allows (at least an eighth of them)
			{
				const char* e = (flags & CFB_LOAD_NO_WALK8) ? "0" : getenv("CFB_WALK8");
				const uint64_t nrows = h.len + 1;
				uint64_t cover = std::min<uint64_t>(nrows, budget() / 8);
				{ const char* f = getenv("CFB_WALK8_ROWS"); if(f) cover = std::min<uint64_t>(nrows, strtoull(f, NULL, 10)); }     // tests: force a partial table
				if(!(e && e[0] == '0') && nrows < (1ull << 40) && cover >= nrows / 8 && cover > 0) {
					void* tab = nullptr;
					CKX(cudaMalloc(&tab, cover * 8 + 16));
					ix->dptrs.push_back(tab); ix->device_bytes += cover * 8;
					int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_build_walk8, kSearchThreads, 0);
					k_build_walk8<<<prop.multiProcessorCount * std::max(occ, 1) * 4, kSearchThreads>>>(v, cover, (uint64_t*)tab);
					CKX(cudaDeviceSynchronize());
					v.walk8 = (const uint64_t*)tab; v.walk8_rows = cover;
					ix->tables.walk8_bytes = cover * 8; ix->tables.walk8_rows = cover;
				}
			}
		}
		{ size_t fb = 0, tb = 0; cudaMemGetInfo(&fb, &tb); ix->tables.total_bytes = ix->device_bytes; ix->tables.free_bytes_after_load = fb; }
		// host copies of the big arrays are no longer needed once uploaded
		std::vector<uint8_t>().swap(ix->h.sides);
	}
	*out = ix;
	return CFB_OK;
}
#undef CKX
extern "C" int cfb_index_load(const char* basename, int device, cfb_index** out) { return cfb_index_load_ex(basename, device, 0u, out); }
extern "C" void cfb_index_free(cfb_index* ix) {
	if(!ix) return;
	if(ix->device >= 0) { cudaSetDevice(ix->device); for(size_t i = 0; i < ix->dptrs.size(); i++) cudaFree(ix->dptrs[i]); }
	delete ix;
}
extern "C" int cfb_index_get_info(const cfb_index* ix, cfb_index_info* o) {
	if(!ix || !o) return fail(CFB_EINVAL, "null argument");
	const HostIndex& h = ix->h;
	o->len = h.len; o->num_sides = h.num_sides; o->n_seqs = h.seq_taxid.size(); o->n_tax_nodes = h.nodes.size();
	o->n_boundaries = h.brow.size(); o->line_rate = h.line_rate; o->off_rate = h.off_rate; o->ftab_chars = h.ftab_chars;
	o->sample_bytes = h.wide_sample ? 4 : 2; o->compressed = h.compressed ? 1 : 0; o->device = ix->device; o->device_bytes = ix->device_bytes;
	return CFB_OK;
}
extern "C" int cfb_index_get_tables(const cfb_index* ix, cfb_index_tables* o) {
	if(!ix || !o) return fail(CFB_EINVAL, "null argument");
	*o = ix->tables;
	return CFB_OK;
}
extern "C" const char* cfb_index_seq_name(const cfb_index* ix, uint32_t s) { return (ix && s < ix->h.seq_name.size()) ? ix->h.seq_name[s].c_str() : NULL; }
extern "C" uint64_t cfb_index_seq_taxid(const cfb_index* ix, uint32_t s) { return (ix && s < ix->h.seq_taxid.size()) ? ix->h.seq_taxid[s] : 0; }
extern "C" int cfb_index_tax_node(const cfb_index* ix, uint64_t taxid, uint64_t* parent, int* rank, int* leaf) {
	const TaxNode* n = ix ? ix->h.find_node(taxid) : NULL;
	if(!n) return 0;
	if(parent) *parent = n->parent; if(rank) *rank = n->rank; if(leaf) *leaf = n->leaf;
	return 1;
}
// internal accessor for the host driver (cf_host.cpp); not part of the public C ABI
extern "C" const cfb::HostIndex* cfb_index_host(const cfb_index* ix) { return ix ? &ix->h : NULL; }
extern "C" int cfb_index_node_taxids(const cfb_index* ix, uint64_t* out, uint64_t cap) {
	if(!ix || !out || cap < ix->h.nodes.size()) return fail(CFB_EINVAL, "cfb_index_node_taxids: buffer too small");
	for(size_t i = 0; i < ix->h.nodes.size(); i++) out[i] = ix->h.nodes[i].taxid;
	return CFB_OK;
}
extern "C" void cfb_params_default(cfb_params* p) {
	if(!p) return;
	memset(p, 0, sizeof *p); p->khits = 5; p->min_hitlen = 22; p->tree_traverse = 1; p->class_rank_slot = 0;
}

// ---------------------------------------------------------------------------------------
static const int kSlots = 17;  // 16 pipelined slots (two waves of 8 sub-batches keep the copy engines busy across batch boundaries) + 1 for resident batches

struct Slot {
	cudaStream_t st = nullptr;
	cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	// inputs
	HBuf<uint8_t> h_bases; HBuf<uint64_t> h_off; HBuf<uint32_t> h_len; HBuf<uint8_t> h_flags;
	DBuf<uint8_t> d_bases; DBuf<uint64_t> d_off; DBuf<uint32_t> d_len; DBuf<uint8_t> d_flags;
	HBuf<uint64_t> h_words; DBuf<uint64_t> d_words, d_npos, d_woff; DBuf<uint32_t> d_wlen;      // packed input (cfb_batch_packed)
	// work
	DBuf<uint64_t> pk; DBuf<uint32_t> nm;
	DBuf<HitRec> hits; DBuf<uint32_t> nhits; DBuf<HitRec> regen; DBuf<uint32_t> regen_n; uint64_t regen_slots = 0; uint32_t full_cap = 0; DBuf<uint32_t> nrows; DBuf<uint64_t> row_off; DBuf<uint64_t> bsum;
	DBuf<uint64_t> rows; DBuf<uint32_t> ids; DBuf<Entry> entries; DBuf<TaxCnt> tcs; DBuf<OutRec> sparse;
	DBuf<uint32_t> nout; DBuf<uint64_t> out_off; DBuf<OutRec> dense; DBuf<uint32_t> rec_off32;
	DBuf<unsigned long long> scal;    // [0] search task ctr (u32 used), [1] resolve ctr, [2] overflow, [3] total rows, [4] total recs
	HBuf<unsigned long long> h_scal;
	HBuf<OutRec> h_recs; HBuf<uint32_t> h_rec_off;
	DBuf<unsigned long long> cnt;     // this batch's per-taxon counters (record path), added to the context's totals at wait time
	bool folded = false, is_text = false, commit_pending = false;
	// batch bookkeeping
	BatchView bv; uint64_t n_units = 0, n_bases = 0; uint32_t maxlen = 0, cap = 0; uint64_t rows_cap = 0, dense_cap = 0;
	bool pending = false, reran = false;
	bool want_host = false; uint64_t d2h_recs = 0;     // records already copied to h_recs by the speculative D2H queued behind the kernels
	void release() {
		h_bases.release(); h_off.release(); h_len.release(); h_flags.release(); d_bases.release(); d_off.release(); d_len.release(); d_flags.release();
		h_words.release(); d_words.release(); d_npos.release(); d_woff.release(); d_wlen.release();
		pk.release(); nm.release(); hits.release(); nhits.release(); regen.release(); regen_n.release(); nrows.release(); row_off.release(); bsum.release(); rows.release(); ids.release(); entries.release(); tcs.release();
		sparse.release(); nout.release(); out_off.release(); dense.release(); rec_off32.release(); scal.release(); h_scal.release(); h_recs.release(); h_rec_off.release(); cnt.release();
		for(int i = 0; i < 6; i++) if(ev[i]) cudaEventDestroy(ev[i]);
		if(st) cudaStreamDestroy(st);
	}
};

struct cfb_dbatch { int slot; uint64_t n_units; };
struct TextCtx;                       // cf_text.cuh
static void text_release(cfb_ctx*);
static void comm_release(cfb_ctx*);   // cf_multi.cuh

// Per-taxon counters of a context, on the device (SpeciesMetrics::addSpeciesCounts, aln_sink.h:142-172): for every taxid
// the report can mention -- tree nodes, sequence taxids, 0 and 1 -- {numReads, numUniqueReads, reads whose single
// best row reached the maximum score}.  Both the text operator (k_fmt_plan) and the record-level path (k_fold_counts,
// when cfb_ctx_count_records is on) add to `total` when a batch is collected; cfb_counts_allreduce sums `total` over the
// communicator's ranks into `global` (NCCL, ncclUint64, ncclSum).
struct CountsCtx {
	bool ready = false;
	std::vector<uint64_t> h_taxid; DBuf<uint64_t> d_taxid; uint32_t n = 0;
	DBuf<unsigned long long> total, global; bool reduced = false;
	void release() { d_taxid.release(); total.release(); global.release(); }
};

struct cfb_ctx {
	const cfb_index* ix = nullptr;
	IndexView view; Params prm;
	DBuf<uint8_t> d_excl; DBuf<uint64_t> d_host;
	Slot slots[kSlots];
	Counters* d_ctr = nullptr; int count = 0;        // CFB_COUNT: 1 = reference operation counters, 2 = the product's own load requests
	uint64_t launches = 0;
	int search_blocks = 0, resolve_blocks = 0, group = 1; int resolve_mode = 2;   // 0 = 8-lane sides, 1 = thread/blocks, 2 = 4-lane rank16
	cfb_dbatch resident; bool resident_used = false;
	double rec_ratio = 2.0;       // records per unit seen so far (sizes the speculative D2H)
	uint64_t rows_cap0 = 0;       // CFB_ROWS_CAP: initial row-buffer capacity (tests force the grow-and-re-run path with it)
	TextCtx* text = nullptr;
	CountsCtx cnt; bool fold_records = false;
	uint32_t jump_w = 4; bool keep_short = false;      // CFB_KEEP_SHORT=1: store every hit (A/B and tests)
	uint64_t regen_lists = 0, regen_tasks = 0;   // lists regenerated / strand lists searched so far (CFB_REGEN_STATS=1 prints them when the context goes)
	uint64_t regen_slots0 = 0;    // CFB_REGEN_SLOTS: initial capacity of the list-regeneration buffer (tests force the grow-and-re-run path with it)
	void* comm = nullptr; int comm_rank = 0, comm_size = 1; cudaStream_t comm_st = nullptr;      // NCCL communicator (cf_multi.cuh)
};

// every tree node whose ancestor chain contains a listed id (Classifier ctor classifier.h:157-201)
static void expand_taxids(const HostIndex& h, const uint64_t* ids, uint64_t n, std::set<uint64_t>& out) {
	if(n == 0 || !ids) return;
	for(size_t i = 0; i < h.nodes.size(); i++) {
		uint64_t t = h.nodes[i].taxid;
		for(;;) {
			bool found = false;
			for(uint64_t k = 0; k < n; k++) if(ids[k] == t) { found = true; break; }
			if(found) { out.insert(h.nodes[i].taxid); break; }
			const TaxNode* nd = h.find_node(t);
			if(!nd || nd->parent == t) break;
			t = nd->parent;
		}
	}
}

extern "C" void cfb_ctx_destroy(cfb_ctx* c) {
	if(!c) return;
	if(c->ix && c->ix->device >= 0) cudaSetDevice(c->ix->device);
	if(getenv("CFB_REGEN_STATS") && c->regen_tasks)
		fprintf(stderr, "[cfb] strand lists regenerated by k_prep: %llu of %llu (%.3f %%)\n", (unsigned long long)c->regen_lists,
		        (unsigned long long)c->regen_tasks, 100.0 * (double)c->regen_lists / (double)c->regen_tasks);
	text_release(c);
	comm_release(c);
	c->cnt.release();
	for(int i = 0; i < kSlots; i++) c->slots[i].release();
	c->d_excl.release(); c->d_host.release();
	if(c->d_ctr) cudaFree(c->d_ctr);
	delete c;
}

extern "C" int cfb_ctx_create(const cfb_index* ix, const cfb_params* p, cfb_ctx** out) {
	if(!ix || !p || !out) return fail(CFB_EINVAL, "cfb_ctx_create: null argument");
	if(ix->device < 0) return fail(CFB_ENODEV, "index was loaded host-only; classification needs a CUDA device (no CPU fallback)");
	if(p->khits < 1) return fail(CFB_EINVAL, "khits must be >= 1");
	CK(cudaSetDevice(ix->device));
	cfb_ctx* c = new cfb_ctx();
	c->ix = ix; c->view = ix->view;
	const HostIndex& h = ix->h;
	Params& q = c->prm;
	q.khits = (uint32_t)p->khits;
	q.min_hitlen = (uint32_t)(p->min_hitlen < 15 ? 15 : p->min_hitlen);
	q.ihits = (uint32_t)std::max(p->khits, 5) * (h.compressed ? 4u : 40u);       // ReportingParams aln_sink.h:580-588
	q.increment = (2 * q.min_hitlen <= 33) ? 10 : (2 * q.min_hitlen - 33);        // classifier.h:226
	q.tree_traverse = p->tree_traverse ? 1 : 0;
	q.class_rank_slot = (uint32_t)(p->class_rank_slot & 0xff);
	std::set<uint64_t> host, excl;
	expand_taxids(h, p->host_taxids, p->n_host_taxids, host);
	expand_taxids(h, p->excluded_taxids, p->n_excluded_taxids, excl);
	#define CKC(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { cfb_ctx_destroy(c); return fail(CFB_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while(0)
	if(!excl.empty()) {
		std::vector<uint8_t> fl(h.seq_taxid.size(), 0);
		for(size_t i = 0; i < fl.size(); i++) fl[i] = excl.count(h.seq_taxid[i]) ? 1 : 0;
		CKC(c->d_excl.ensure(fl.size())); CKC(cudaMemcpy(c->d_excl.p, fl.data(), fl.size(), cudaMemcpyHostToDevice));
		c->view.seq_excluded = c->d_excl.p;
	}
	if(!host.empty()) {
		std::vector<uint64_t> hv(host.begin(), host.end());
		CKC(c->d_host.ensure(hv.size())); CKC(cudaMemcpy(c->d_host.p, hv.data(), hv.size() * 8, cudaMemcpyHostToDevice));
		c->view.host_taxids = c->d_host.p; c->view.n_host = (uint32_t)hv.size();
	}
	for(int i = 0; i < kSlots; i++) {
		CKC(cudaStreamCreateWithFlags(&c->slots[i].st, cudaStreamNonBlocking));
		for(int e = 0; e < 6; e++) CKC(cudaEventCreate(&c->slots[i].ev[e]));
		CKC(c->slots[i].scal.ensure(8)); CKC(c->slots[i].h_scal.ensure(8));
	}
	CKC(cudaMalloc((void**)&c->d_ctr, sizeof(Counters))); CKC(cudaMemset(c->d_ctr, 0, sizeof(Counters)));
	{   // random 32-byte sector gathers: do not let L2 over-fetch neighbouring sectors from HBM
		const char* e = getenv("CFB_L2_FETCH"); const size_t g = e ? (size_t)atoi(e) : 32;
		if(g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
		const char* pe = getenv("CFB_L2_PERSIST");
		if(pe && pe[0] == '1') {   // keep the 8.4 MB ftab resident in L2 (hit by every partial search)
			cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 16u << 20);
			for(int i = 0; i < kSlots; i++) {
				cudaStreamAttrValue av; memset(&av, 0, sizeof av);
				av.accessPolicyWindow.base_ptr = (void*)c->view.ftab; av.accessPolicyWindow.num_bytes = h.ftab.size() * 8;
				av.accessPolicyWindow.hitRatio = 1.0f; av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
				cudaStreamSetAttribute(c->slots[i].st, cudaStreamAttributeAccessPolicyWindow, &av);
			}
		}
		cudaGetLastError();
	}
	int occ = 0;
	{ const char* g = getenv("CFB_GROUP"); if(g) { const int v = atoi(g); if(v == 1 || v == 2 || v == 4 || v == 8 || v == 16) c->group = v; } }
	if(c->group != 1 && !c->view.sides) { cfb_ctx_destroy(c); return fail(CFB_EINVAL, "CFB_GROUP=%d needs the sides resident (CFB_KEEP_SIDES=1)", c->group); }
	CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_kernel(c->group, false), kSearchThreads, 0));
	c->search_blocks = ix->sm_count * std::max(occ, 1);
	{ const char* g = getenv("CFB_RESOLVE"); if(g) c->resolve_mode = atoi(g); else if(c->view.rtab16 || c->view.rtab32) c->resolve_mode = 3; }
	if(c->resolve_mode == 3 && !(c->view.rtab16 || c->view.rtab32)) c->resolve_mode = 2;
	if(c->resolve_mode == 1 && !c->view.blocks) c->resolve_mode = 2;
	if(c->resolve_mode == 0 && !c->view.sides) c->resolve_mode = 2;
	if(c->resolve_mode >= 2) CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resolve_c<false, false>, kSearchThreads, 0));
	else if(c->resolve_mode == 1) CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resolve_t<false>, kSearchThreads, 0));
	else CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_resolve<false>, kSearchThreads, 0));
	c->resolve_blocks = ix->sm_count * std::max(occ, 1);
	{ const char* rc0 = getenv("CFB_ROWS_CAP"); if(rc0) c->rows_cap0 = strtoull(rc0, NULL, 10); }
	{ const char* ks = getenv("CFB_KEEP_SHORT"); c->keep_short = ks && ks[0] == '1'; }
	{ const char* rs = getenv("CFB_REGEN_SLOTS"); if(rs) c->regen_slots0 = strtoull(rs, NULL, 10); }
	{ const char* jw = getenv("CFB_JUMP_W"); c->jump_w = jw ? (uint32_t)std::min(std::max(atoi(jw), 1), 8) : 4u; }      // measured on the bench workload: 1: 3.57, 2: 3.50, 3: 3.46, 4: 3.43 ms per 2 M reads (profiles/r02_ab.txt)
	const char* cnt = getenv("CFB_COUNT");
	c->count = cnt ? (cnt[0] == '1' ? 1 : (cnt[0] == '2' ? 2 : 0)) : 0;
	#undef CKC
	*out = c;
	return CFB_OK;
}
extern "C" int cfb_ctx_slots(const cfb_ctx*) { return kSlots - 1; }   // last slot is reserved for resident batches
extern "C" int cfb_ctx_kernel_launches(const cfb_ctx* c, uint64_t* n) { if(!c || !n) return CFB_EINVAL; *n = c->launches; return CFB_OK; }

// ---------------------------------------------------------------------------------------
// per-taxon counters of the record-level path
// ---------------------------------------------------------------------------------------
static int counts_init(cfb_ctx* c) {
	CountsCtx& k = c->cnt;
	if(k.ready) return CFB_OK;
	const HostIndex& h = c->ix->h;
	std::set<uint64_t> sp; sp.insert(0); sp.insert(1);
	for(size_t i = 0; i < h.nodes.size(); i++) sp.insert(h.nodes[i].taxid);
	sp.insert(h.seq_taxid.begin(), h.seq_taxid.end());
	k.h_taxid.assign(sp.begin(), sp.end()); k.n = (uint32_t)k.h_taxid.size();
	CK(k.d_taxid.ensure(k.n + 1)); CK(cudaMemcpy(k.d_taxid.p, k.h_taxid.data(), (size_t)k.n * 8, cudaMemcpyHostToDevice));
	CK(k.total.ensure(3ull * k.n)); CK(cudaMemset(k.total.p, 0, 3ull * k.n * 8));
	CK(k.global.ensure(3ull * k.n)); CK(cudaMemset(k.global.p, 0, 3ull * k.n * 8));
	k.ready = true;
	return CFB_OK;
}

struct FoldArgs {
	const uint64_t* sp_taxid; uint32_t n_sp;
	const uint32_t* rec_off; const OutRec* recs; uint32_t n_units; int32_t n_mates; uint32_t khits;
	const uint32_t* len[2]; const uint8_t* flags;
	unsigned long long* sp;        // 3 * n_sp: numReads | numUniqueReads | observed singletons
};
__device__ __forceinline__ int find_slot(const uint64_t* a, uint32_t n, uint64_t key) {
	uint32_t lo = 0, hi = n;
	while(lo < hi) { const uint32_t mid = (lo + hi) >> 1; if(a[mid] < key) lo = mid + 1; else hi = mid; }
	return (lo < n && a[lo] == key) ? (int)lo : -1;
}
// thread per unit: what AlnSinkWrap::finishRead -> SpeciesMetrics::addSpeciesCounts accumulate for the unit's reported
// assignments (aln_sink.h:142-172,1861-1927): the records with the best score, at most khits of them (hit-map order; a tie
// of more than khits records is resolved by the per-read RNG in the reference and in the text operator -- it only arises
// under --host-taxids).  A unit without records counts as taxid 0, as its "unclassified" row does.
__global__ void __launch_bounds__(128) k_fold_counts(const FoldArgs a) {
	const uint32_t u = blockIdx.x * 128 + threadIdx.x;
	const bool live = u < a.n_units;
	int first_slot = -1; uint32_t num = 1; bool qualifies = false;
	if(live) {
		const uint32_t r0 = a.rec_off[u], r1 = a.rec_off[u + 1];
		if(r1 == r0) { first_slot = find_slot(a.sp_taxid, a.n_sp, 0); qualifies = true; }
		else {
			uint32_t best = 0, ties = 0;
			for(uint32_t k = r0; k < r1; k++) { const uint32_t sc = a.recs[k].score; if(sc > best || k == r0) { best = sc; ties = 1; } else if(sc == best) ties++; }
			num = ties < a.khits ? ties : a.khits;
			const uint32_t fl = a.flags ? a.flags[u] : 3u;
			int64_t max_score = 0;
			if(fl & 1u) { const int64_t L = a.len[0][u]; max_score += L > 15 ? (L - 15) * (L - 15) : 0; }
			if(a.n_mates == 2 && (fl & 2u)) { const int64_t L = a.len[1][u]; max_score += L > 15 ? (L - 15) * (L - 15) : 0; }
			qualifies = (int64_t)best >= max_score;
			uint32_t taken = 0;
			for(uint32_t k = r0; k < r1 && taken < num; k++) {
				if(a.recs[k].score != best) continue;
				const int slot = find_slot(a.sp_taxid, a.n_sp, a.recs[k].taxid);
				if(taken == 0) first_slot = slot; else if(slot >= 0) atomicAdd(a.sp + slot, 1ull);
				taken++;
			}
		}
	}
	// first assignment of every unit: warp-aggregated (dominant taxa would serialise per-lane atomics)
	const int key = live ? first_slot : -1;
	const uint32_t peers = __match_any_sync(0xffffffffu, key);
	if(key >= 0) {
		const uint32_t uniq = __popc(__ballot_sync(peers, num == 1) & peers);
		const uint32_t obs = __popc(__ballot_sync(peers, num == 1 && qualifies) & peers);
		if((uint32_t)(__ffs(peers) - 1) == (threadIdx.x & 31u)) {
			atomicAdd(a.sp + key, (unsigned long long)__popc(peers));
			if(uniq) atomicAdd(a.sp + a.n_sp + key, (unsigned long long)uniq);
			if(obs) atomicAdd(a.sp + 2ull * a.n_sp + key, (unsigned long long)obs);
		}
	}
}
__global__ void k_cnt_commit(const unsigned long long* slot_sp, unsigned long long* total, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) { const unsigned long long v = slot_sp[i]; if(v) atomicAdd(total + i, v); }
}

// Validate + stage a batch into the slot's pinned buffers and enqueue H2D copies.
static int stage_batch(cfb_ctx* c, Slot& s, const cfb_batch* b) {
	if(!b || b->n_mates < 1 || b->n_mates > 2 || !b->bases || !b->off[0] || !b->len[0] || (b->n_mates == 2 && (!b->off[1] || !b->len[1])))
		return fail(CFB_EINVAL, "malformed cfb_batch");
	if(b->n_units >= (1ull << 30)) return fail(CFB_EINVAL, "batch too large (n_units must be < 2^30)");
	const uint64_t n = b->n_units; const int nm = b->n_mates;
	uint32_t maxlen = 0;
	for(int m = 0; m < nm; m++) {       // branch-free validation pass (vectorises); the offender is looked up only when there is one
		const uint64_t* O = b->off[m]; const uint32_t* L = b->len[m]; const uint64_t nb = b->n_bases; uint32_t mx = 0; uint64_t far = 0;
		for(uint64_t i = 0; i < n; i++) { const uint32_t l = L[i]; mx = l > mx ? l : mx; }
		for(uint64_t i = 0; i < n; i++) { const uint64_t e = O[i] + (uint64_t)L[i]; far = e > far ? e : far; }
		if(far > nb) { for(uint64_t i = 0; i < n; i++) if(O[i] + L[i] > nb) return fail(CFB_EINVAL, "unit %llu mate %d exceeds n_bases", (unsigned long long)i, m + 1); }
		maxlen = std::max(maxlen, mx);
	}
	if(maxlen > 60000) return fail(CFB_EINVAL, "read longer than 60000 bases");
	CK(s.d_bases.ensure(b->n_bases + 16)); CK(s.d_off.ensure(n * nm)); CK(s.d_len.ensure(n * nm)); CK(s.d_flags.ensure(n));
	// Caller arrays that already live in pinned memory (cfb_host_alloc) are DMA'd from where they are and
	// must stay untouched until cfb_classify_wait; pageable arrays are staged through pinned buffers first.
	auto pinned = [](const void* p) -> bool {
		cudaPointerAttributes at;
		if(cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
		return at.type == cudaMemoryTypeHost;
	};
	const uint8_t* src_bases = b->bases;
	if(!pinned(b->bases)) { CK(s.h_bases.ensure(b->n_bases)); memcpy(s.h_bases.p, b->bases, b->n_bases); src_bases = s.h_bases.p; }
	CK(cudaMemcpyAsync(s.d_bases.p, src_bases, b->n_bases, cudaMemcpyHostToDevice, s.st));
	CK(s.h_off.ensure(n * nm)); CK(s.h_len.ensure(n * nm)); CK(s.h_flags.ensure(n));
	for(int m = 0; m < nm; m++) {
		const uint64_t* so = b->off[m]; const uint32_t* sl = b->len[m];
		if(!pinned(so)) { memcpy(s.h_off.p + m * n, so, n * 8); so = s.h_off.p + m * n; }
		if(!pinned(sl)) { memcpy(s.h_len.p + m * n, sl, n * 4); sl = s.h_len.p + m * n; }
		CK(cudaMemcpyAsync(s.d_off.p + m * n, so, n * 8, cudaMemcpyHostToDevice, s.st));
		CK(cudaMemcpyAsync(s.d_len.p + m * n, sl, n * 4, cudaMemcpyHostToDevice, s.st));
	}
	if(b->flags && pinned(b->flags)) CK(cudaMemcpyAsync(s.d_flags.p, b->flags, n, cudaMemcpyHostToDevice, s.st));
	else { if(b->flags) memcpy(s.h_flags.p, b->flags, n); else memset(s.h_flags.p, 3, n); CK(cudaMemcpyAsync(s.d_flags.p, s.h_flags.p, n, cudaMemcpyHostToDevice, s.st)); }
	s.bv.bases = s.d_bases.p; s.bv.flags = s.d_flags.p; s.bv.n_units = (uint32_t)n; s.bv.n_mates = nm;
	for(int m = 0; m < 2; m++) { s.bv.off[m] = m < nm ? s.d_off.p + m * n : nullptr; s.bv.len[m] = m < nm ? s.d_len.p + m * n : nullptr; }
	s.n_units = n; s.n_bases = b->n_bases; s.maxlen = maxlen;
	(void)c;
	return CFB_OK;
}

// ---------------------------------------------------------------------------------------
// Packed input (cfb_batch_packed): 2 bits per base + a sparse list of N positions instead of 1 byte per base, lengths
// instead of offsets -- about a third of the host->device bytes of cfb_batch.  The device expands it into the byte form
// the per-unit kernels read: every mate gets a 32-byte aligned slot of ceil(len/32)*32 bytes, so byte address =
// 32 * word index + base in word, for the unpack kernel and for the N list alike.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wlen(const uint32_t* len, uint64_t n, uint32_t* wlen) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) wlen[i] = (len[i] + 31u) >> 5;
}
// thread per (mate slot, word k): 32 codes of one packed word -> 32 bytes
__global__ void __launch_bounds__(256) k_unpack(const uint64_t* __restrict__ words, const uint64_t* __restrict__ woff, uint64_t wbase, const uint32_t* __restrict__ len,
                                                uint64_t n, uint32_t W, uint8_t* bases, uint64_t* off) {
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= n * W) return;
	const uint64_t i = t / W; const uint32_t k = (uint32_t)(t - i * W);
	const uint64_t w0 = wbase + woff[i];
	if(k == 0) off[i] = w0 * 32;
	if(k * 32 >= len[i]) return;
	const uint64_t w = __ldg(words + w0 + k);
	uint32_t o[8];
	#pragma unroll
	for(int q = 0; q < 8; q++) {          // 4 codes -> 4 bytes
		const uint32_t b = (uint32_t)(w >> (8 * q)) & 0xffu;
		o[q] = (b & 3u) | ((b & 0xcu) << 6) | ((b & 0x30u) << 12) | ((b & 0xc0u) << 18);
	}
	uint4* dst = reinterpret_cast<uint4*>(bases + (w0 + k) * 32);
	dst[0] = make_uint4(o[0], o[1], o[2], o[3]); dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}
__global__ void __launch_bounds__(256) k_set_n(const uint64_t* npos, uint64_t n, uint8_t* bases, uint64_t limit) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n) { const uint64_t p = npos[i]; if(p < limit) bases[p] = 4; }
}

static int stage_batch_packed(cfb_ctx* c, Slot& s, const cfb_batch_packed* b) {
	if(!b || b->n_mates < 1 || b->n_mates > 2 || !b->words || !b->len[0] || (b->n_mates == 2 && !b->len[1]) || (b->n_n && !b->n_pos))
		return fail(CFB_EINVAL, "malformed cfb_batch_packed");
	if(b->n_units >= (1ull << 30)) return fail(CFB_EINVAL, "batch too large (n_units must be < 2^30)");
	const uint64_t n = b->n_units; const int nm = b->n_mates;
	// one branch-free pass per mate over the lengths (it vectorises: this runs on the submitting thread, once per batch)
	uint32_t maxlen = 0; uint64_t need_words = 0, mate_words[2] = {0, 0};
	for(int m = 0; m < nm; m++) {
		const uint32_t* L = b->len[m]; uint32_t mx = 0; uint64_t w = 0;
		for(uint64_t i = 0; i < n; i++) { const uint32_t l = L[i]; mx = l > mx ? l : mx; w += (l + 31u) >> 5; }
		maxlen = std::max(maxlen, mx); mate_words[m] = w; need_words += w;
	}
	if(need_words != b->n_words) return fail(CFB_EINVAL, "cfb_batch_packed: n_words is %llu, the lengths need %llu", (unsigned long long)b->n_words, (unsigned long long)need_words);
	if(maxlen > 60000) return fail(CFB_EINVAL, "read longer than 60000 bases");
	const uint64_t scan_blocks = (n + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer);
	CK(s.d_bases.ensure(b->n_words * 32 + 64)); CK(s.d_off.ensure(n * nm + 1)); CK(s.d_len.ensure(n * nm)); CK(s.d_flags.ensure(n));
	CK(s.d_words.ensure(b->n_words + 1)); CK(s.d_npos.ensure(b->n_n + 1)); CK(s.d_wlen.ensure(n + 1)); CK(s.d_woff.ensure(n + 2)); CK(s.bsum.ensure(scan_blocks + 1));
	auto pinned = [](const void* p) -> bool {
		cudaPointerAttributes at;
		if(cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
		return at.type == cudaMemoryTypeHost;
	};
	const uint64_t* src_words = b->words;
	if(!pinned(b->words)) { CK(s.h_words.ensure(b->n_words)); memcpy(s.h_words.p, b->words, b->n_words * 8); src_words = s.h_words.p; }
	CK(cudaMemcpyAsync(s.d_words.p, src_words, b->n_words * 8, cudaMemcpyHostToDevice, s.st));
	CK(s.h_len.ensure(n * nm)); CK(s.h_flags.ensure(n));
	for(int m = 0; m < nm; m++) {
		const uint32_t* sl = b->len[m];
		if(!pinned(sl)) { memcpy(s.h_len.p + m * n, sl, n * 4); sl = s.h_len.p + m * n; }
		CK(cudaMemcpyAsync(s.d_len.p + m * n, sl, n * 4, cudaMemcpyHostToDevice, s.st));
	}
	if(b->n_n) {
		const uint64_t* sp = b->n_pos;
		if(!pinned(sp)) { CK(s.h_off.ensure(b->n_n)); memcpy(s.h_off.p, sp, b->n_n * 8); sp = s.h_off.p; }
		CK(cudaMemcpyAsync(s.d_npos.p, sp, b->n_n * 8, cudaMemcpyHostToDevice, s.st));
	}
	if(b->flags && pinned(b->flags)) CK(cudaMemcpyAsync(s.d_flags.p, b->flags, n, cudaMemcpyHostToDevice, s.st));
	else { if(b->flags) memcpy(s.h_flags.p, b->flags, n); else memset(s.h_flags.p, 3, n); CK(cudaMemcpyAsync(s.d_flags.p, s.h_flags.p, n, cudaMemcpyHostToDevice, s.st)); }
	// expand on the device
	const uint32_t W = (maxlen + 31) / 32;
	uint64_t wbase = 0;
	CK(cudaMemsetAsync(s.scal.p + 5, 0, 8, s.st));
	for(int m = 0; m < nm && n; m++) {
		k_wlen<<<(unsigned)((n + 255) / 256), 256, 0, s.st>>>(s.d_len.p + m * n, n, s.d_wlen.p);
		k_scan_sums<<<(unsigned)scan_blocks, kScanBlock, 0, s.st>>>(s.d_wlen.p, n, s.bsum.p);
		k_scan_top<<<1, 1024, 0, s.st>>>(s.bsum.p, scan_blocks, (uint64_t*)(s.scal.p + 5));
		k_scan_apply<<<(unsigned)scan_blocks, kScanBlock, 0, s.st>>>(s.d_wlen.p, n, s.bsum.p, (const uint64_t*)(s.scal.p + 5), s.d_woff.p);
		if(W) k_unpack<<<(unsigned)((n * W + 255) / 256), 256, 0, s.st>>>(s.d_words.p, s.d_woff.p, wbase, s.d_len.p + m * n, n, W, s.d_bases.p, s.d_off.p + m * n);
		c->launches += 5;
		wbase += mate_words[m];      // mate 2 starts after all of mate 1
	}
	if(b->n_n) { k_set_n<<<(unsigned)((b->n_n + 255) / 256), 256, 0, s.st>>>(s.d_npos.p, b->n_n, s.d_bases.p, b->n_words * 32); c->launches++; }
	CK(cudaGetLastError());
	s.bv.bases = s.d_bases.p; s.bv.flags = s.d_flags.p; s.bv.n_units = (uint32_t)n; s.bv.n_mates = nm;
	for(int m = 0; m < 2; m++) { s.bv.off[m] = m < nm ? s.d_off.p + m * n : nullptr; s.bv.len[m] = m < nm ? s.d_len.p + m * n : nullptr; }
	s.n_units = n; s.n_bases = b->n_words * 32; s.maxlen = maxlen;
	return CFB_OK;
}

// Enqueue all kernels of one batch on the slot's stream.  stage: 0 = from search, 1 = from k_rows
// (after a rows-capacity overflow; the hit lists are already post-processed and sorted).
static int enqueue_kernels(cfb_ctx* c, Slot& s, int stage, bool time_it) {
	const uint64_t n = s.n_units; const int nm = s.bv.n_mates;
	const uint64_t ntasks = n * nm * 2;
	const uint32_t ublocks = (uint32_t)((n + 127) / 128);
	const uint64_t scan_blocks = (n + kScanBlock * kScanPer - 1) / (kScanBlock * kScanPer);
	if(n == 0) return CFB_OK;
	// which search kernel runs decides what it stores: the thread-per-walk kernels keep only hits of >= kLongLen bases when
	// min_hitlen allows it (see kListRegen), the generic / A-B kernels and small min_hitlen keep every hit
	int variant = c->group;
	if(variant == 1) { if(s.maxlen > 320) variant = 16; else if(s.maxlen > 160) variant = 100; else if(s.maxlen > 128) variant = 101; }   // longer reads: wider register window / windowed kernel
	const bool pooled = variant == 1 || variant == 100 || variant == 101;
	const bool keep_short = !pooled || c->prm.min_hitlen < kLongLen || c->keep_short;
	if(stage == 0) {
		if(s.cap == 0) {
			s.full_cap = s.maxlen / 4 + 8;      // >= #Ns allowed by the N filter (0.15 len) + len/10 + slack
			s.cap = keep_short ? s.full_cap : s.maxlen / kLongLen + 2;       // hits of >= 22 bases do not overlap
		}
		CK(s.hits.ensure(ntasks * s.cap)); CK(s.nhits.ensure(ntasks));
		if(!keep_short) {
			s.regen_slots = std::max<uint64_t>(s.regen_slots, c->regen_slots0 ? c->regen_slots0 : std::max<uint64_t>(ntasks / 32, 1024));
			CK(s.regen.ensure(s.regen_slots * s.full_cap)); CK(s.regen_n.ensure(s.regen_slots));
			CK(cudaMemsetAsync(s.scal.p + 6, 0, sizeof(unsigned long long), s.st));
		}
		const uint32_t W = (s.maxlen + 31) / 32 + 1;
		CK(s.pk.ensure(ntasks * W + 2)); CK(s.nm.ensure(ntasks * W + 2)); CK(s.nrows.ensure(n)); CK(s.row_off.ensure(n + 1));
		CK(s.bsum.ensure(scan_blocks + 1)); CK(s.nout.ensure(n)); CK(s.out_off.ensure(n + 1)); CK(s.rec_off32.ensure(n + 1));
		s.rows_cap = std::max<uint64_t>(s.rows_cap, c->rows_cap0 ? c->rows_cap0 : std::max<uint64_t>(n * 12, 4096));
	}
	CK(s.rows.ensure(s.rows_cap)); CK(s.ids.ensure(s.rows_cap)); CK(s.entries.ensure(s.rows_cap)); CK(s.tcs.ensure(s.rows_cap)); CK(s.sparse.ensure(s.rows_cap));
	s.dense_cap = s.rows_cap; CK(s.dense.ensure(s.dense_cap));
	CK(cudaMemsetAsync(s.scal.p, 0, 4 * sizeof(unsigned long long), s.st));   // task counters, overflow flag, row allocator; [4] is rewritten by the scan
	if(time_it) CK(cudaEventRecord(s.ev[0], s.st));
	Counters* ctr = c->count ? c->d_ctr : nullptr;
	if(c->count && stage == 0) CK(cudaMemsetAsync(c->d_ctr, 0, sizeof(Counters), s.st));
	UnitArgs ua; ua.v = c->view; ua.p = c->prm; ua.b = s.bv; ua.hits = s.hits.p; ua.nhits = s.nhits.p; ua.cap = s.cap;
	ua.nrows = s.nrows.p; ua.row_off = s.row_off.p; ua.row_total = s.scal.p + 3; ua.rows = s.rows.p; ua.ids = s.ids.p; ua.rows_cap = s.rows_cap;
	ua.entries = s.entries.p; ua.tcs = s.tcs.p; ua.recs_sparse = s.sparse.p; ua.nout = s.nout.p;
	ua.overflow = (unsigned int*)(s.scal.p + 2); ua.ctr = ctr;
	ua.regen = s.regen.p; ua.regen_n = s.regen_n.p; ua.regen_ctr = s.scal.p + 6; ua.regen_slots = s.regen_slots; ua.full_cap = s.full_cap; ua.keep_short = keep_short ? 1u : 0u;
	if(stage == 0) {
		SearchArgs sa; sa.v = c->view; sa.p = c->prm; sa.b = s.bv; sa.hits = s.hits.p; sa.nhits = s.nhits.p; sa.cap = s.cap;
		const uint32_t W = (s.maxlen + 31) / 32 + 1;
		sa.pk = s.pk.p; sa.nm = s.nm.p; sa.W = W; sa.jump_w = c->jump_w; sa.keep_short = keep_short ? 1u : 0u;
		{ PackArgs pa; pa.b = s.bv; pa.pk = s.pk.p; pa.nm = s.nm.p; pa.W = W;
		  k_pack<<<(unsigned)((ntasks * W + 127) / 128), 128, 0, s.st>>>(pa); c->launches++; }
		sa.task_ctr = (unsigned int*)(s.scal.p + 0); sa.task_ctr64 = s.scal.p + 0; sa.ntasks = (uint32_t)ntasks; sa.overflow = (unsigned int*)(s.scal.p + 2); sa.ctr = ctr;
		const int lanes = (variant == 16 || pooled) ? 1 : variant;
		int occ = 1;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_kernel(variant, c->count), kSearchThreads, 0));
		const uint64_t resident = (uint64_t)c->ix->sm_count * (uint64_t)std::max(occ, 1);
		const uint64_t per_block = kSearchThreads / lanes;
		const uint64_t groups = resident * per_block;
		uint64_t chunk = ntasks / (groups * 8); sa.chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(chunk, 1), 16);
		if(pooled) {   // one atomic per warp refill of >= 32 tasks
			const uint64_t warps = resident * (kSearchThreads / 32);
			sa.chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(ntasks / (warps * 4), 32), 256);
		}
		const int blocks = (int)std::min<uint64_t>(resident, (ntasks + per_block - 1) / per_block);
		search_kernel(variant, c->count)<<<blocks, kSearchThreads, 0, s.st>>>(sa);
		c->launches++;
		if(time_it) CK(cudaEventRecord(s.ev[1], s.st));
		k_prep<8, false><<<ublocks, 128, 0, s.st>>>(ua); c->launches++;          // <= 64 registers (measured best of 72 / 64 / 40)
	} else {
		if(time_it) CK(cudaEventRecord(s.ev[1], s.st));
		k_prep<8, true><<<ublocks, 128, 0, s.st>>>(ua); c->launches++;
	}
	if(time_it) CK(cudaEventRecord(s.ev[2], s.st));
	ResolveArgs ra; ra.v = c->view; ra.rows = s.rows.p; ra.ids = s.ids.p; ra.ids16 = nullptr; ra.total = (const uint64_t*)(s.scal.p + 3); ra.rows_cap = s.rows_cap;
	ra.task_ctr = s.scal.p + 1; ra.chunk = c->resolve_mode >= 2 ? 64 : (c->resolve_mode == 1 ? 128 : 4); ra.ctr = ctr;
	if(c->resolve_mode == 3 && c->count != 1) k_lookup<<<c->ix->sm_count * 8, 256, 0, s.st>>>(ra);
	else if(c->resolve_mode >= 2) { if(c->count) k_resolve_c<true, false><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); else k_resolve_c<false, false><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); }
	else if(c->resolve_mode == 1) { if(c->count) k_resolve_t<true><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); else k_resolve_t<false><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); }
	else { if(c->count) k_resolve<true><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); else k_resolve<false><<<c->resolve_blocks, kSearchThreads, 0, s.st>>>(ra); }
	c->launches++;
	if(time_it) CK(cudaEventRecord(s.ev[3], s.st));
	k_score<12><<<ublocks, 128, 0, s.st>>>(ua); c->launches++;           // <= 40 registers: occupancy beats the few spills (measured)
	k_scan_sums<<<(unsigned)scan_blocks, kScanBlock, 0, s.st>>>(s.nout.p, n, s.bsum.p);
	k_scan_top<<<1, 1024, 0, s.st>>>(s.bsum.p, scan_blocks, (uint64_t*)(s.scal.p + 4));
	k_scan_apply<<<(unsigned)scan_blocks, kScanBlock, 0, s.st>>>(s.nout.p, n, s.bsum.p, (const uint64_t*)(s.scal.p + 4), s.out_off.p);
	k_compact<<<(unsigned)((n + 1 + 127) / 128), 128, 0, s.st>>>((uint32_t)n, s.row_off.p, s.out_off.p, s.sparse.p, s.dense.p, s.rec_off32.p, s.dense_cap, (unsigned int*)(s.scal.p + 2));
	c->launches += 4;
	s.folded = false;
	if(c->fold_records && !s.is_text) {      // this batch's per-taxon counters, on the device, from the records just written
		const uint32_t nsp = c->cnt.n;
		CK(s.cnt.ensure(3ull * nsp)); CK(cudaMemsetAsync(s.cnt.p, 0, 3ull * nsp * 8, s.st));
		FoldArgs fa; fa.sp_taxid = c->cnt.d_taxid.p; fa.n_sp = nsp; fa.rec_off = s.rec_off32.p; fa.recs = s.dense.p; fa.n_units = (uint32_t)n; fa.n_mates = nm;
		fa.khits = c->prm.khits; fa.len[0] = s.bv.len[0]; fa.len[1] = s.bv.len[1]; fa.flags = s.bv.flags; fa.sp = s.cnt.p;
		k_fold_counts<<<ublocks, 128, 0, s.st>>>(fa); c->launches++;
		s.folded = true;
	}
	if(time_it) CK(cudaEventRecord(s.ev[4], s.st));
	s.d2h_recs = 0;
	if(s.want_host) {
		// results go home behind the kernels without waiting for the host to learn their size: record offsets
		// exactly, records for the count the previous batches suggest (finish_batch fetches a remainder if any)
		const uint64_t guess = std::min<uint64_t>(s.dense_cap, (uint64_t)((double)n * c->rec_ratio * 1.05) + 4096);
		CK(s.h_recs.ensure(guess + 1)); CK(s.h_rec_off.ensure(n + 1));
		CK(cudaMemcpyAsync(s.h_rec_off.p, s.rec_off32.p, (n + 1) * 4, cudaMemcpyDeviceToHost, s.st));
		CK(cudaMemcpyAsync(s.h_recs.p, s.dense.p, guess * sizeof(OutRec), cudaMemcpyDeviceToHost, s.st));
		s.d2h_recs = guess;
	}
	CK(cudaMemcpyAsync(s.h_scal.p, s.scal.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.st));
	CK(cudaGetLastError());
	return CFB_OK;
}

// Wait for the kernels, handle capacity overflows by re-running the affected stages, then D2H.
static int finish_batch(cfb_ctx* c, Slot& s, bool time_it, bool to_host, cfb_result* out) {
	if(s.n_units == 0) { if(out) { out->n_units = 0; out->n_recs = 0; out->rec_off = nullptr; out->recs = nullptr; } return CFB_OK; }
	for(int attempt = 0; attempt < 8; attempt++) {
		CK(cudaStreamSynchronize(s.st));
		const unsigned ovf = (unsigned)(s.h_scal.p[2] & 0xffffffffu);
		const uint64_t total_rows = s.h_scal.p[3];
		if(ovf == 4) {            // more lists needed regeneration than the side buffer holds: h_scal[6] tells how many
			s.regen_slots = s.h_scal.p[6] + s.h_scal.p[6] / 4 + 1024; s.reran = true;
			int rc = enqueue_kernels(c, s, 0, time_it); if(rc) return rc;
			continue;
		}
		if(ovf == 1) {            // hit-list capacity: only possible when the caller's flags bypass the N filter
			s.cap = s.maxlen + 2; s.full_cap = s.maxlen + 2; s.reran = true;
			int rc = enqueue_kernels(c, s, 0, time_it); if(rc) return rc;
			continue;
		}
		if(total_rows > s.rows_cap) {
			s.rows_cap = total_rows + total_rows / 4 + 1024; s.reran = true;
			int rc = enqueue_kernels(c, s, 1, time_it); if(rc) return rc;
			continue;
		}
		if(ovf != 0) return fail(CFB_ECUDA, "internal capacity error %u", ovf);
		break;
	}
	const uint64_t nrec = s.h_scal.p[4];
	c->regen_lists += s.h_scal.p[6]; c->regen_tasks += (uint64_t)s.n_units * (uint64_t)s.bv.n_mates * 2;
	if(s.folded) {        // the batch is final: add its counters to the context's totals (stream order keeps this ahead of any read)
		const uint32_t n3 = 3 * c->cnt.n;
		k_cnt_commit<<<(n3 + 255) / 256, 256, 0, s.st>>>(s.cnt.p, c->cnt.total.p, n3); c->launches++;
		CK(cudaEventRecord(s.ev[5], s.st)); s.commit_pending = true;       // cfb_counts_allreduce orders itself behind this event, without a host-side wait
		s.folded = false; c->cnt.reduced = false;
	}
	if(to_host) {
		if(s.n_units) c->rec_ratio = std::max(c->rec_ratio * 0.98, (double)nrec / (double)s.n_units);
		if(!s.want_host || nrec > s.d2h_recs) {       // not (fully) covered by the speculative copy
			const uint64_t have = s.want_host ? s.d2h_recs : 0;
			if(nrec + 1 > s.h_recs.cap) {             // grow, keeping nothing: copy everything again
				CK(s.h_recs.ensure(nrec + 1));
				if(nrec) CK(cudaMemcpyAsync(s.h_recs.p, s.dense.p, nrec * sizeof(OutRec), cudaMemcpyDeviceToHost, s.st));
			} else if(nrec > have) CK(cudaMemcpyAsync(s.h_recs.p + have, s.dense.p + have, (nrec - have) * sizeof(OutRec), cudaMemcpyDeviceToHost, s.st));
			if(!s.want_host) { CK(s.h_rec_off.ensure(s.n_units + 1)); CK(cudaMemcpyAsync(s.h_rec_off.p, s.rec_off32.p, (s.n_units + 1) * 4, cudaMemcpyDeviceToHost, s.st)); }
			CK(cudaStreamSynchronize(s.st));
		}
	}
	if(out) { out->n_units = s.n_units; out->n_recs = nrec; out->rec_off = s.h_rec_off.p; out->recs = reinterpret_cast<const cfb_rec*>(s.h_recs.p); }
	return CFB_OK;
}

extern "C" int cfb_classify_submit(cfb_ctx* c, int slot, const cfb_batch* b) {
	if(!c || slot < 0 || slot >= kSlots - 1) return fail(CFB_EINVAL, "bad ctx/slot");
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[slot];
	if(s.pending) return fail(CFB_EINVAL, "slot %d still has an un-waited batch", slot);
	int rc = stage_batch(c, s, b); if(rc) return rc;
	s.cap = 0; s.want_host = true; s.is_text = false;
	rc = enqueue_kernels(c, s, 0, false); if(rc) return rc;
	s.pending = true;
	return CFB_OK;
}
extern "C" int cfb_classify_submit_packed(cfb_ctx* c, int slot, const cfb_batch_packed* b) {
	if(!c || slot < 0 || slot >= kSlots - 1) return fail(CFB_EINVAL, "bad ctx/slot");
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[slot];
	if(s.pending) return fail(CFB_EINVAL, "slot %d still has an un-waited batch", slot);
	int rc = stage_batch_packed(c, s, b); if(rc) return rc;
	s.cap = 0; s.want_host = true; s.is_text = false;
	rc = enqueue_kernels(c, s, 0, false); if(rc) return rc;
	s.pending = true;
	return CFB_OK;
}
// Host helper: the packed form of a cfb_batch (single pass, one thread; callers that parse reads themselves can emit
// the packed form directly).  Returns CFB_EINVAL when a capacity is too small; n_words / n_n always receive the sizes needed.
extern "C" int cfb_pack_batch(const cfb_batch* in, uint64_t* words, uint64_t words_cap, uint64_t* n_pos, uint64_t npos_cap, uint64_t* n_words, uint64_t* n_n) {
	if(!in || !n_words || !n_n || in->n_mates < 1 || in->n_mates > 2) return fail(CFB_EINVAL, "cfb_pack_batch: bad argument");
	uint64_t w = 0, nn = 0; bool fits = true;
	for(int m = 0; m < in->n_mates; m++) for(uint64_t i = 0; i < in->n_units; i++) {
		const uint8_t* p = in->bases + in->off[m][i]; const uint32_t len = in->len[m][i];
		for(uint32_t k = 0; k < len; k += 32) {
			uint64_t v = 0; const uint32_t cnt = std::min<uint32_t>(32, len - k);
			for(uint32_t j = 0; j < cnt; j++) {
				const uint8_t c = p[k + j];
				if(c > 3) { if(n_pos && nn < npos_cap) n_pos[nn] = (w << 5) | j; else fits = false; nn++; }
				else v |= (uint64_t)c << (2 * j);
			}
			if(words && w < words_cap) words[w] = v; else fits = false;
			w++;
		}
	}
	*n_words = w; *n_n = nn;
	return fits ? CFB_OK : fail(CFB_EINVAL, "cfb_pack_batch: buffers too small (%llu words, %llu N positions needed)", (unsigned long long)w, (unsigned long long)nn);
}
extern "C" int cfb_classify_wait(cfb_ctx* c, int slot, cfb_result* out) {
	if(!c || slot < 0 || slot >= kSlots - 1 || !out) return fail(CFB_EINVAL, "bad ctx/slot");
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[slot];
	if(!s.pending) return fail(CFB_EINVAL, "slot %d has no submitted batch", slot);
	s.pending = false;
	return finish_batch(c, s, false, true, out);
}
extern "C" int cfb_classify_batch(cfb_ctx* c, const cfb_batch* b, cfb_result* out) {
	int rc = cfb_classify_submit(c, 0, b); if(rc) return rc;
	return cfb_classify_wait(c, 0, out);
}

extern "C" int cfb_batch_upload(cfb_ctx* c, const cfb_batch* b, cfb_dbatch** out) {
	if(!c || !out) return fail(CFB_EINVAL, "null argument");
	if(c->resident_used) return fail(CFB_EINVAL, "only one resident batch per ctx");
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[kSlots - 1];
	int rc = stage_batch(c, s, b); if(rc) return rc;
	CK(cudaStreamSynchronize(s.st));
	s.cap = 0; s.is_text = false;
	c->resident.slot = kSlots - 1; c->resident.n_units = s.n_units; c->resident_used = true;
	*out = &c->resident;
	return CFB_OK;
}
extern "C" void cfb_dbatch_free(cfb_ctx* c, cfb_dbatch*) { if(c) c->resident_used = false; }
extern "C" int cfb_classify_resident(cfb_ctx* c, cfb_dbatch* d, float* ms, uint64_t* n_recs) {
	return cfb_classify_resident_range(c, d, 0, d ? d->n_units : 0, ms, n_recs);
}
extern "C" int cfb_classify_resident_range(cfb_ctx* c, cfb_dbatch* d, uint64_t first, uint64_t count, float* ms, uint64_t* n_recs) {
	if(!c || !d) return fail(CFB_EINVAL, "null argument");
	if(first + count > d->n_units) return fail(CFB_EINVAL, "cfb_classify_resident_range: units [%llu, %llu) exceed the uploaded batch", (unsigned long long)first, (unsigned long long)(first + count));
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[d->slot];
	// a window of the uploaded batch: offsets are absolute into the uploaded bases, so only the per-unit arrays shift
	const uint64_t N = d->n_units; const int nmates = s.bv.n_mates;
	s.bv.flags = s.d_flags.p + first; s.bv.n_units = (uint32_t)count; s.n_units = count;
	for(int m = 0; m < nmates; m++) { s.bv.off[m] = s.d_off.p + m * N + first; s.bv.len[m] = s.d_len.p + m * N + first; }
	int rc = enqueue_kernels(c, s, 0, true); if(rc) return rc;
	cfb_result r;
	rc = finish_batch(c, s, true, false, &r); if(rc) return rc;
	if(n_recs) *n_recs = r.n_recs;
	if(ms) {
		CK(cudaEventSynchronize(s.ev[4]));
		CK(cudaEventElapsedTime(&ms[0], s.ev[0], s.ev[1])); CK(cudaEventElapsedTime(&ms[1], s.ev[1], s.ev[2]));
		CK(cudaEventElapsedTime(&ms[2], s.ev[2], s.ev[3])); CK(cudaEventElapsedTime(&ms[3], s.ev[3], s.ev[4]));
		CK(cudaEventElapsedTime(&ms[4], s.ev[0], s.ev[4]));
	}
	return CFB_OK;
}
extern "C" int cfb_resident_result(cfb_ctx* c, cfb_result* out) {
	if(!c || !out || !c->resident_used) return fail(CFB_EINVAL, "no resident batch");
	CK(cudaSetDevice(c->ix->device));
	Slot& s = c->slots[kSlots - 1];
	return finish_batch(c, s, false, true, out);
}
extern "C" int cfb_ctx_counters(cfb_ctx* c, uint64_t out[8]) {
	if(!c || !out) return fail(CFB_EINVAL, "null argument");
	CK(cudaSetDevice(c->ix->device));
	Counters h; CK(cudaMemcpy(&h, c->d_ctr, sizeof h, cudaMemcpyDeviceToHost));
	out[0] = h.units; out[1] = h.partial_searches; out[2] = h.ftab_probes; out[3] = h.sides_search;
	out[4] = h.walk_steps; out[5] = h.rows_resolved; out[6] = h.lf_steps; out[7] = h.ext_searches;
	return CFB_OK;
}

// pinned for every device of the process (several contexts on different GPUs may DMA from the same buffer pool)
extern "C" void* cfb_host_alloc(size_t bytes) { void* p = nullptr; if(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
extern "C" int cfb_device_count(void) { int n = 0; if(cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
extern "C" void cfb_host_free(void* p) { if(p) cudaFreeHost(p); }

extern "C" int cfb_test_lf(const cfb_index* ix, const uint64_t* rows, const uint8_t* chars, uint64_t n, uint64_t* out) {
	if(!ix || ix->device < 0) return fail(CFB_ENODEV, "no device");
	if(!ix->view.sides) return fail(CFB_EINVAL, "the sides are not resident on this replica (CFB_KEEP_SIDES=1 keeps them)");
	CK(cudaSetDevice(ix->device));
	uint64_t *dr = nullptr, *dout = nullptr; uint8_t* dc = nullptr;
	CK(cudaMalloc(&dr, n * 8 + 8)); CK(cudaMalloc(&dout, n * 8 + 8)); CK(cudaMalloc(&dc, n + 8));
	CK(cudaMemcpy(dr, rows, n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dc, chars, n, cudaMemcpyHostToDevice));
	k_test_lf<<<64, 128>>>(ix->view, dr, dc, n, dout);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost));
	cudaFree(dr); cudaFree(dout); cudaFree(dc);
	return CFB_OK;
}
extern "C" int cfb_test_resolve(const cfb_index* ix, const uint64_t* rows, uint64_t n, uint32_t* out) {
	if(!ix || ix->device < 0) return fail(CFB_ENODEV, "no device");
	if(!ix->view.sides) return fail(CFB_EINVAL, "the sides are not resident on this replica (CFB_KEEP_SIDES=1 keeps them)");
	CK(cudaSetDevice(ix->device));
	uint64_t* dr = nullptr; uint32_t* dout = nullptr; unsigned long long* sc = nullptr;
	CK(cudaMalloc(&dr, n * 8 + 8)); CK(cudaMalloc(&dout, n * 4 + 8)); CK(cudaMalloc(&sc, 16));
	CK(cudaMemcpy(dr, rows, n * 8, cudaMemcpyHostToDevice));
	unsigned long long init[2] = {0ull, (unsigned long long)n};
	CK(cudaMemcpy(sc, init, 16, cudaMemcpyHostToDevice));
	ResolveArgs ra; ra.v = ix->view; ra.rows = dr; ra.ids = dout; ra.ids16 = nullptr; ra.total = (const uint64_t*)(sc + 1); ra.rows_cap = n; ra.task_ctr = sc; ra.chunk = 2; ra.ctr = nullptr;
	k_resolve<false><<<32, kSearchThreads>>>(ra);
	CK(cudaDeviceSynchronize());
	CK(cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost));
	cudaFree(dr); cudaFree(dout); cudaFree(sc);
	return CFB_OK;
}

#include "cf_text.cuh"
#include "cf_multi.cuh"
