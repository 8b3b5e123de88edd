// micro-benchmark: what device allocation costs on this box (cudaMalloc / cudaFree vs size; VMM map of chunks)
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main()
{
	cudaFree(0);
	for (int round = 0; round < 2; round++)
		for (size_t mb : {1, 16, 64, 256, 1024, 2048}) {
			void *p;
			double t0 = now();
			cudaMalloc(&p, mb << 20);
			double t1 = now();
			cudaMemset(p, 0, mb << 20);
			cudaDeviceSynchronize();
			double t2 = now();
			cudaFree(p);
			double t3 = now();
			printf("round %d  %5zu MB: malloc %.3f ms, memset %.3f ms, free %.3f ms\n", round, mb, t1 - t0, t2 - t1, t3 - t2);
		}
	// many small allocations alive at once (the engine's ~60 arrays)
	{
		std::vector<void *> v(64);
		double t0 = now();
		for (auto &p : v) cudaMalloc(&p, 8 << 20);
		double t1 = now();
		for (auto &p : v) cudaFree(p);
		double t2 = now();
		printf("64 x 8 MB: malloc %.3f ms, free %.3f ms\n", t1 - t0, t2 - t1);
	}
	// VMM: reserve 8 GB of address space, map 64 MB chunks one by one
	{
		CUmemAllocationProp prop = {};
		prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
		prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
		prop.location.id = 0;
		size_t gran = 0;
		cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
		printf("VMM granularity %zu\n", gran);
		CUdeviceptr base;
		double t0 = now();
		CUresult r = cuMemAddressReserve(&base, (size_t)8 << 30, 0, 0, 0);
		double t1 = now();
		printf("reserve 8 GB VA: %.3f ms (rc %d)\n", t1 - t0, (int)r);
		const size_t chunk = (size_t)64 << 20;
		CUmemAccessDesc acc = {};
		acc.location = prop.location;
		acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
		std::vector<CUmemGenericAllocationHandle> hs;
		for (int k = 0; k < 16; k++) {
			CUmemGenericAllocationHandle h;
			double a = now();
			cuMemCreate(&h, chunk, &prop, 0);
			double b = now();
			cuMemMap(base + k * chunk, chunk, 0, h, 0);
			double c = now();
			cuMemSetAccess(base + k * chunk, chunk, &acc, 1);
			double d = now();
			hs.push_back(h);
			if (k < 3 || k == 15) printf("chunk %d (64 MB): create %.3f map %.3f access %.3f ms\n", k, b - a, c - b, d - c);
		}
		cudaMemset((void *)base, 1, 16 * chunk);
		printf("memset of the mapped GB: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	}
	return 0;
}
