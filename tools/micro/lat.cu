// Latency micro-benchmarks behind the cost model of the tail kernels (profiles/README.md):
// __syncthreads at 1024 threads, a 1024-thread block scan, barrier.cluster at 8 CTAs, a dependent
// global-load hop (L2 hit), an atomicExch round trip.  Build: nvcc -arch=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
__device__ __forceinline__ u64 gt() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { u32 n = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += n; }
	return v;
}
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32 *ws, u32 &total)
{
	const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	u32 incl = warp_incl_scan(v);
	if (lane == 31) ws[wid] = incl;
	__syncthreads();
	if (wid == 0) { u32 s = lane < nw ? ws[lane] : 0; u32 si = warp_incl_scan(s); ws[lane] = si - s; if (lane == 31) ws[32] = si; }
	__syncthreads();
	u32 r = incl - v + ws[wid];
	total = ws[32];
	__syncthreads();
	return r;
}
__global__ void __launch_bounds__(1024, 1) k(u32 *chain, u32 *scratch, u64 *out, int reps)
{
	__shared__ u32 ws[33];
	u32 rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	const bool t0 = rank == 0 && threadIdx.x == 0;
	u64 a, b;
	// (a) __syncthreads
	__syncthreads(); a = gt();
	for (int i = 0; i < reps; i++) __syncthreads();
	b = gt(); if (t0) out[0] = b - a;
	// (b) block scan
	u32 acc = threadIdx.x, tot;
	__syncthreads(); a = gt();
	for (int i = 0; i < reps; i++) acc = block_excl_scan(acc & 3, ws, tot);
	b = gt(); if (t0) out[1] = b - a;
	// (c) cluster barrier
	a = gt();
	for (int i = 0; i < reps; i++) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	b = gt(); if (t0) out[2] = b - a;
	// (d) dependent loads (pointer chase through a 64 MB permutation: L2 hits after the warm-up pass)
	if (t0) {
		u32 p = 0;
		for (int i = 0; i < reps; i++) p = chain[p];
		a = gt();
		p = 0;
		for (int i = 0; i < reps; i++) p = chain[p];
		b = gt(); out[3] = b - a; scratch[1] = p;
		// (e) atomicExch round trips
		a = gt();
		u32 q = 0;
		for (int i = 0; i < reps; i++) q = atomicExch(&scratch[64 + (q & 1023)], i);
		b = gt(); out[4] = b - a; scratch[2] = q;
		// (f) store then cluster-visible: store + __threadfence
		a = gt();
		for (int i = 0; i < reps; i++) { scratch[4096 + i] = i; __threadfence(); }
		b = gt(); out[5] = b - a;
	}
	out[8 + rank] = acc + tot;
}
int main()
{
	const int reps = 64;
	const size_t n = 16u << 20;                    // 64 MB of u32
	u32 *h = (u32 *)malloc(n * 4);
	// a long stride permutation: p -> (p + 1048583 * 16) mod n keeps successive hops in different lines
	for (size_t i = 0; i < n; i++) h[i] = (u32)((i + 16777259ull % n * 0 + 1048583ull * 16) % n);
	u32 *chain, *scratch; u64 *out;
	cudaMalloc(&chain, n * 4); cudaMalloc(&scratch, 1 << 20); cudaMalloc(&out, 64 * 8);
	cudaMemcpy(chain, h, n * 4, cudaMemcpyHostToDevice); cudaMemset(scratch, 0, 1 << 20);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(8); cfg.blockDim = dim3(1024);
	cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
	cfg.attrs = at; cfg.numAttrs = 1;
	for (int it = 0; it < 3; it++) {
		cudaLaunchKernelEx(&cfg, k, chain, scratch, out, reps);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
		u64 o[8]; cudaMemcpy(o, out, sizeof o, cudaMemcpyDeviceToHost);
		printf("run %d (ns per op): syncthreads=%.0f block_scan=%.0f cluster_barrier=%.0f load_hop=%.0f atomic_exch=%.0f store+fence=%.0f\n", it,
		       o[0] / (double)reps, o[1] / (double)reps, o[2] / (double)reps, o[3] / (double)reps, o[4] / (double)reps, o[5] / (double)reps);
	}
	return 0;
}
