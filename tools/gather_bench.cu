// gather_bench.cu -- what the HBM system of this GPU delivers for *independent random* gathers of
// 32 / 64 / 128 bytes from a multi-GB array (the access pattern of the FM-index walk), as opposed to the
// streaming-copy peak in MEASURED_PEAKS.json.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
template <int BYTES, int ILP, bool DEP>
__global__ void k_gather(const uint4* a, uint64_t nunits, uint64_t iters, uint64_t* out) {
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t acc = 0, h = mix(tid);
	for(uint64_t it = 0; it < iters; it++) {
		uint4 v[ILP * (BYTES / 16)];
		#pragma unroll
		for(int k = 0; k < ILP; k++) {
			h = mix(h + k + (DEP ? acc : 0));                   // DEP: next address depends on loaded data (pointer chase)
			const uint64_t u = h % nunits;
			#pragma unroll
			for(int q = 0; q < BYTES / 16; q++) v[k * (BYTES / 16) + q] = __ldg(a + u * (BYTES / 16) + q);
		}
		#pragma unroll
		for(int k = 0; k < ILP * (BYTES / 16); k++) acc += v[k].x ^ v[k].w;
	}
	if(acc == 0x1234567) out[0] = acc;
}
template <int BYTES, int ILP, bool DEP> void run(const uint4* a, uint64_t bytes, uint64_t* out, int blocks, int threads, const char* tag) {
	const uint64_t nunits = bytes / BYTES, iters = 64;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	k_gather<BYTES, ILP, DEP><<<blocks, threads>>>(a, nunits, 4, out); cudaDeviceSynchronize();
	cudaEventRecord(e0); k_gather<BYTES, ILP, DEP><<<blocks, threads>>>(a, nunits, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	const double n = (double)blocks * threads * iters * ILP;
	printf("%-28s %3d B x ilp %d  blocks %5d: %7.2f G gathers/s  %8.1f GB/s useful  (%.2f ms)\n", tag, BYTES, ILP, blocks, n / ms / 1e6, n * BYTES / ms / 1e6, ms);
}
int main() {
	const uint64_t bytes = 12ull << 30;
	uint4* a; uint64_t* out; cudaMalloc(&a, bytes); cudaMalloc(&out, 8); cudaMemset(a, 1, bytes);
	int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	for(int occ : {8, 16}) {
		const int blocks = sms * occ;
		run<32, 1, true>(a, bytes, out, blocks, 128, "dependent chain");
		run<32, 1, false>(a, bytes, out, blocks, 128, "independent");
		run<32, 4, false>(a, bytes, out, blocks, 128, "independent");
		run<64, 1, true>(a, bytes, out, blocks, 128, "dependent chain");
		run<64, 4, false>(a, bytes, out, blocks, 128, "independent");
		run<128, 1, true>(a, bytes, out, blocks, 128, "dependent chain");
		run<128, 4, false>(a, bytes, out, blocks, 128, "independent");
	}
	return 0;
}
