// Micro-benchmark (GPU box): the slot pipeline of cs_multi.cu reduced to its CUDA calls, to find what keeps the input copy of a
// resubmitted slot from starting while result copies of other slots are queued.
//   slot: [ev0] H2D 150 MB  [memset ctrl] [ev1] spin kernel (K ms, full grid) [D2D 4 B] [evk]  | host sees evk -> [evc] D2H 225 MB in 5 copies [evd]
// variants: bit0 = no memset / D2D copies, bit1 = D2H as ONE copy, bit2 = H2D through a copy KERNEL reading mapped host memory,
//           bit3 = D2H through a copy KERNEL writing mapped host memory, bit4 = result copies on a second stream of the slot (after an event)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <thread>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) ; }
__global__ void kcopy(const uint4 *src, uint4 *dst, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i]; }
struct SlotS { cudaStream_t st, so; cudaEvent_t e0, e1, ek, ec, ed; char *hin, *hout, *din, *dout; int *ctrl; int state; int batch; };
int main(int argc, char **argv)
{
	const int variant = argc > 1 ? atoi(argv[1]) : 0, NSLOT = argc > 2 ? atoi(argv[2]) : 3, NB = 10;
	const size_t NI = 150u << 20, NO = 225u << 20;
	std::vector<SlotS> S(NSLOT);
	cudaEvent_t base; CK(cudaEventCreate(&base));
	for (auto &s : S) {
		CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s.so, cudaStreamNonBlocking));
		CK(cudaEventCreate(&s.e0)); CK(cudaEventCreate(&s.e1)); CK(cudaEventCreate(&s.ek)); CK(cudaEventCreate(&s.ec)); CK(cudaEventCreate(&s.ed));
		CK(cudaHostAlloc(&s.hin, NI, cudaHostAllocMapped)); CK(cudaHostAlloc(&s.hout, NO, cudaHostAllocMapped)); CK(cudaMalloc(&s.din, NI)); CK(cudaMalloc(&s.dout, NO)); CK(cudaMalloc(&s.ctrl, 256));
		memset(s.hin, 1, NI); s.state = 0;
	}
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(base, S[0].st)); CK(cudaEventSynchronize(base));
	auto t0 = std::chrono::steady_clock::now();
	auto hms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	int next = 0, done = 0, copied = 0;
	std::vector<float> tl(NB * 8, 0.f);
	auto submit = [&](int b) {
		SlotS &s = S[b % NSLOT]; s.batch = b; tl[b * 8 + 5] = (float)hms();
		CK(cudaEventRecord(s.e0, s.st));
		if (variant & 4) kcopy<<<148 * 2, 256, 0, s.st>>>((const uint4*)s.hin, (uint4*)s.din, NI / 16);
		else CK(cudaMemcpyAsync(s.din, s.hin, NI, cudaMemcpyHostToDevice, s.st));
		if (!(variant & 1)) CK(cudaMemsetAsync(s.ctrl, 0, 256, s.st));
		CK(cudaEventRecord(s.e1, s.st));
		spin<<<148 * 4, 256, 0, s.st>>>(14000000ll);
		if (!(variant & 1)) CK(cudaMemcpyAsync(s.ctrl + 8, s.ctrl, 4, cudaMemcpyDeviceToDevice, s.st));
		CK(cudaEventRecord(s.ek, s.st));
		s.state = 2;
	};
	while (copied < NB) {
		bool prog = false;
		while (next < NB && next - copied < NSLOT) { submit(next++); prog = true; }
		if (done < next) { SlotS &s = S[done % NSLOT];
			if (cudaEventQuery(s.ek) == cudaSuccess) {
				tl[done * 8 + 6] = (float)hms();
				cudaStream_t so = (variant & 16) ? s.so : s.st;
				if (variant & 16) CK(cudaStreamWaitEvent(so, s.ek, 0));
				CK(cudaEventRecord(s.ec, so));
				if (variant & 8) kcopy<<<148 * 2, 256, 0, so>>>((const uint4*)s.dout, (uint4*)s.hout, NO / 16);
				else if (variant & 2) CK(cudaMemcpyAsync(s.hout, s.dout, NO, cudaMemcpyDeviceToHost, so));
				else for (int k = 0; k < 5; ++k) CK(cudaMemcpyAsync(s.hout + k * (NO / 5), s.dout + k * (NO / 5), NO / 5, cudaMemcpyDeviceToHost, so));
				CK(cudaEventRecord(s.ed, so));
				++done; prog = true;
			} else cudaGetLastError(); }
		if (copied < done) { SlotS &s = S[copied % NSLOT];
			if (cudaEventQuery(s.ed) == cudaSuccess) {
				cudaEvent_t e[5] = { s.e0, s.e1, s.ek, s.ec, s.ed };
				for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&tl[copied * 8 + k], base, e[k]);
				tl[copied * 8 + 7] = (float)hms(); ++copied; prog = true;
			} else cudaGetLastError(); }
		if (!prog) std::this_thread::sleep_for(std::chrono::microseconds(20));
	}
	printf("variant %d slots %d: total %.1f ms for %d batches (kernel alone %.1f ms each)\n", variant, NSLOT, hms(), NB, 14000000.0 / 1.965e6);
	for (int b = 0; b < NB; ++b) printf("  %d: in[%.1f %.1f] kern[..%.1f] out[%.1f %.1f] host: submit %.1f seen %.1f done %.1f\n", b, tl[b*8], tl[b*8+1], tl[b*8+2], tl[b*8+3], tl[b*8+4], tl[b*8+5], tl[b*8+6], tl[b*8+7]);
	return 0;
}
