// Micro-benchmark (GPU box): do host-to-device and device-to-host copies on different streams overlap, from one host thread,
// with page-locked memory from cudaMallocHost vs cudaHostRegister, with and without a long kernel on a third stream?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/copy_overlap scripts/micro/copy_overlap.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void spin(long long cycles, int *sink) { long long t0 = clock64(); while (clock64() - t0 < cycles) ; if (sink && threadIdx.x == 9999) *sink = 1; }

int main(int argc, char **argv)
{
	const size_t N = 256u << 20;
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	printf("asyncEngineCount %d  CUDA_DEVICE_MAX_CONNECTIONS=%s\n", p.asyncEngineCount, getenv("CUDA_DEVICE_MAX_CONNECTIONS") ? getenv("CUDA_DEVICE_MAX_CONNECTIONS") : "(unset)");
	char *d1, *d2; CK(cudaMalloc(&d1, N)); CK(cudaMalloc(&d2, N));
	for (int mode = 0; mode < 2; ++mode) {
		char *h1, *h2;
		if (mode == 0) { CK(cudaMallocHost(&h1, N)); CK(cudaMallocHost(&h2, N)); }
		else { h1 = (char*)aligned_alloc(4096, N); h2 = (char*)aligned_alloc(4096, N); memset(h1, 1, N); memset(h2, 1, N); CK(cudaHostRegister(h1, N, cudaHostRegisterDefault)); CK(cudaHostRegister(h2, N, cudaHostRegisterDefault)); }
		cudaStream_t s1, s2, s3; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
		cudaEvent_t e[8]; for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&e[i]));
		for (int withk = 0; withk < 2; ++withk)
			for (int split = 0; split < 2; ++split) {
				for (int rep = 0; rep < 2; ++rep) {
					CK(cudaDeviceSynchronize());
					CK(cudaEventRecord(e[0], s1));
					if (withk) spin<<<148 * 4, 256, 0, s3>>>(40000000ll, nullptr);   // ~20 ms
					CK(cudaEventRecord(e[1], s1));
					if (split) for (int k = 0; k < 5; ++k) CK(cudaMemcpyAsync(h1 + k * (N / 5), d1 + k * (N / 5), N / 5, cudaMemcpyDeviceToHost, s1));
					else CK(cudaMemcpyAsync(h1, d1, N, cudaMemcpyDeviceToHost, s1));
					CK(cudaEventRecord(e[2], s1));
					CK(cudaEventRecord(e[3], s2));
					CK(cudaMemcpyAsync(d2, h2, N, cudaMemcpyHostToDevice, s2));
					CK(cudaEventRecord(e[4], s2));
					CK(cudaDeviceSynchronize());
					float a, b, c, d;
					cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[0], e[2]); cudaEventElapsedTime(&c, e[0], e[3]); cudaEventElapsedTime(&d, e[0], e[4]);
					if (rep) printf("%s kernel=%d d2h_in_5=%d : D2H [%.2f, %.2f] ms  H2D [%.2f, %.2f] ms  (256 MiB each: %.1f / %.1f GB/s)\n", mode ? "cudaHostRegister" : "cudaMallocHost  ", withk, split, a, b, c, d,
					                N / ((b - a) * 1e6), N / ((d - c) * 1e6));
				}
			}
		if (mode == 0) { cudaFreeHost(h1); cudaFreeHost(h2); } else { cudaHostUnregister(h1); cudaHostUnregister(h2); free(h1); free(h2); }
	}
	// many streams: does stream i + 8 wait for stream i (hardware queue aliasing)?
	{
		const int NS = 14; cudaStream_t s[NS]; cudaEvent_t b[NS], f[NS], e0; CK(cudaEventCreate(&e0));
		for (int i = 0; i < NS; ++i) { CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking)); CK(cudaEventCreate(&b[i])); CK(cudaEventCreate(&f[i])); }
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0, s[0]));
		for (int i = 0; i < NS; ++i) { CK(cudaEventRecord(b[i], s[i])); spin<<<8, 64, 0, s[i]>>>(10000000ll, nullptr); CK(cudaEventRecord(f[i], s[i])); }
		CK(cudaDeviceSynchronize());
		printf("14 streams, one 5 ms 8-CTA kernel each (start, end ms):");
		for (int i = 0; i < NS; ++i) { float x, y; cudaEventElapsedTime(&x, e0, b[i]); cudaEventElapsedTime(&y, e0, f[i]); printf(" [%.1f %.1f]", x, y); }
		printf("\n");
	}
	return 0;
}
