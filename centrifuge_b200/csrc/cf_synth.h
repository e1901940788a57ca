// cf_synth.h -- deterministic synthetic genomes as a pure function of (seed, sequence, position).
//
// Same recipe as tools/synth.py / SURVEY.md Appendix C (G genera x S species; a genus base is iid
// uniform ACGT, every other species of the genus copies it with per-base substitution probability
// `div`), but counter-based so that any slice of any genome can be produced anywhere -- on the GPU
// for the index builder, and for the read sampler -- without materialising multi-Gbp FASTA files.
#ifndef CF_SYNTH_H_
#define CF_SYNTH_H_
#include <stdint.h>

#ifdef __CUDACC__
#define CFS_HD __host__ __device__ __forceinline__
#else
#define CFS_HD inline
#endif

namespace cfb {

struct SynthSpec { uint32_t genera, species; uint64_t len, seed; uint32_t div_q32; };   // div as fraction of 2^32

CFS_HD uint64_t mix64(uint64_t x) {      // splitmix64 finaliser
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

// base (0..3) of sequence `seq` (= genus * species + s) at position pos
CFS_HD int synth_base(const SynthSpec& sp, uint32_t seq, uint64_t pos) {
	const uint32_t genus = seq / sp.species, s = seq - genus * sp.species;
	const uint64_t h = mix64(sp.seed * 0x100000001B3ull + ((uint64_t)genus << 40) + pos);
	int c = (int)(h & 3);
	if(s != 0) {
		const uint64_t u = mix64((sp.seed ^ 0xA5A5A5A5ull) * 0x100000001B3ull + ((uint64_t)(seq + 1) << 40) + pos);
		if((uint32_t)(u >> 32) < sp.div_q32) c = (c + 1 + (int)((u & 0xffff) % 3)) & 3;
	}
	return c;
}

}  // namespace cfb
#endif
