// On-device FM-index construction (needed here because the headline
// configuration is a 3.1 Gbp reference and the reference's single-threaded bwaidx takes hours).
//
// Produces exactly what bwa_idx_build produces (FM_index/index_main.c:257-325): the BWT of
// T$ = fwd + revcomp(fwd) + '$' with '$' smallest and its row dropped, Occ checkpoints, and the
// suffix array sampled every sa_intv rows with sa[0] = -1 (bwt_cal_sa, bwt.c:62-84) -- but written
// directly in the device layout of cs_device.cuh.
//
// Suffix sort: MSD bucket by the first 2 bases, then an LSD radix sort (cub::DeviceRadixSort, the
// only library primitive used, and only here, off the hot path) of each bucket on the next 30
// bases packed into a 64-bit key; suffixes that still tie after 32 bases are refined 32 bases at
// a time.  Positions past the end read as 'A' and ties are broken by the remaining length, which
// reproduces the "'$' is the smallest symbol" order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include "cs_kernels.cuh"

cs_index_t *cs_index_adopt(int device, uint4 *d_buckets, uint64_t n_buckets, uint64_t *d_sa, uint64_t n_sa, int sa_intv,
                           uint64_t primary, const uint64_t L2[5], uint64_t seq_len, const uint64_t *W, const cs_index_config_t *cfg);
void cs_internal_set_error(int code, const char *msg);

namespace {

thread_local char b_err[512];

#define BCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	snprintf(b_err, sizeof b_err, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); goto fail; } } while (0)

// ---- text: 32 bases per u64, base j of a word at bits 62-2j (so integer order == lexicographic) ----
__global__ void k_pack_text(const uint8_t *fwd, uint64_t L, uint64_t *W, uint64_t n_words)
{
	const uint64_t n = 2 * L;
	for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t v = 0;
		for (int j = 0; j < 32; ++j) {
			uint64_t i = (w << 5) + j;
			uint64_t c = 0;
			if (i < L) c = fwd[i] & 3;
			else if (i < n) c = 3 - (fwd[n - 1 - i] & 3);
			v |= c << (62 - 2 * j);
		}
		W[w] = v;
	}
}

__device__ __forceinline__ uint64_t key_at(const uint64_t *W, uint64_t n, uint64_t pos)
{ // the 32 bases starting at pos, zero ('A') padded past the end; W has two zero words of slack
	if (pos >= n) return 0;
	uint64_t w = pos >> 5; uint32_t s = 2 * ((uint32_t)pos & 31);
	uint64_t k = W[w] << s;
	if (s) k |= W[w + 1] >> (64 - s);
	return k;
}
__device__ __forceinline__ uint32_t base_at(const uint64_t *W, uint64_t i) { return (uint32_t)(W[i >> 5] >> (62 - 2 * (i & 31))) & 3; }

// ---- MSD pass: partition positions by their first two bases ----
#define PB_CHUNK 4096
__global__ void k_prefix_hist(const uint64_t *W, uint64_t n, unsigned long long *hist)
{
	__shared__ unsigned int h[16];
	if (threadIdx.x < 16) h[threadIdx.x] = 0;
	__syncthreads();
	for (uint64_t base = (uint64_t)blockIdx.x * PB_CHUNK; base < n; base += (uint64_t)gridDim.x * PB_CHUNK)
		for (uint32_t j = threadIdx.x; j < PB_CHUNK; j += blockDim.x) {
			uint64_t i = base + j;
			if (i < n) atomicAdd(&h[key_at(W, n, i) >> 60], 1u);
		}
	__syncthreads();
	if (threadIdx.x < 16 && h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

__global__ void k_prefix_scatter(const uint64_t *W, uint64_t n, unsigned long long *cursor, uint64_t *sa)
{
	__shared__ unsigned int h[16];
	__shared__ unsigned long long start[16];
	for (uint64_t base = (uint64_t)blockIdx.x * PB_CHUNK; base < n; base += (uint64_t)gridDim.x * PB_CHUNK) {
		if (threadIdx.x < 16) h[threadIdx.x] = 0;
		__syncthreads();
		for (uint32_t j = threadIdx.x; j < PB_CHUNK; j += blockDim.x) {
			uint64_t i = base + j;
			if (i < n) atomicAdd(&h[key_at(W, n, i) >> 60], 1u);
		}
		__syncthreads();
		if (threadIdx.x < 16) { start[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)h[threadIdx.x]); h[threadIdx.x] = 0; }
		__syncthreads();
		for (uint32_t j = threadIdx.x; j < PB_CHUNK; j += blockDim.x) {
			uint64_t i = base + j;
			if (i < n) {
				uint32_t p = (uint32_t)(key_at(W, n, i) >> 60);
				sa[start[p] + atomicAdd(&h[p], 1u)] = i;
			}
		}
		__syncthreads();
	}
}

__global__ void k_keys(const uint64_t *W, uint64_t n, const uint64_t *pos, uint64_t cnt, uint64_t d, uint64_t *keys)
{
	for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < cnt; j += (uint64_t)gridDim.x * blockDim.x)
		keys[j] = key_at(W, n, pos[j] + d);
}

// ---- ties ----
__global__ void k_count_ties(const uint64_t *keys, uint64_t cnt, unsigned long long *m)
{
	unsigned long long c = 0;
	for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < cnt; j += (uint64_t)gridDim.x * blockDim.x)
		c += (j > 0 && keys[j] == keys[j - 1]) || (j + 1 < cnt && keys[j] == keys[j + 1]);
	if (c) atomicAdd(m, c);
}

__global__ void k_emit_ties(const uint64_t *keys, const uint64_t *vals, uint64_t cnt, uint64_t off, unsigned long long *m,
                            uint64_t *t_slot, uint64_t *t_val, uint8_t *t_head)
{
	for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < cnt; j += (uint64_t)gridDim.x * blockDim.x) {
		bool prev = j > 0 && keys[j] == keys[j - 1], next = j + 1 < cnt && keys[j] == keys[j + 1];
		if (prev || next) {
			unsigned long long t = atomicAdd(m, 1ull);
			t_slot[t] = off + j; t_val[t] = vals[j]; t_head[t] = !prev;
		}
	}
}

struct MaxOp { __device__ __forceinline__ uint64_t operator()(uint64_t a, uint64_t b) const { return a > b ? a : b; } };

__global__ void k_head_slots(const uint64_t *slot, const uint8_t *head, uint64_t m, uint64_t *out)
{
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x)
		out[t] = head[t] ? slot[t] + 1 : 0; // +1 so that slot 0 survives the max-scan
}

__global__ void k_refine_keys(const uint64_t *W, uint64_t n, const uint64_t *val, uint64_t m, uint64_t d, uint64_t *key2, uint32_t *rem,
                              uint32_t *perm, unsigned int *any_short)
{
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t p = val[t] + d;
		key2[t] = key_at(W, n, p);
		long long r = (long long)n - (long long)p;           // bases left at offset d (may be <= 0)
		if (r < 32) atomicOr(any_short, 1u);
		if (r > 32) r = 32;
		if (r < -(1ll << 30)) r = -(1ll << 30);
		rem[t] = (uint32_t)(r + (1ll << 31));
		perm[t] = (uint32_t)t;
	}
}

template <typename T>
__global__ void k_gather(const T *src, const uint32_t *perm, uint64_t m, T *dst)
{
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x)
		dst[t] = src[perm[t]];
}

// after the refinement sort: write the re-ordered suffixes back and flag what is still tied
__global__ void k_refine_apply(const uint64_t *slot, const uint64_t *grp_s, const uint64_t *key_s, const uint32_t *rem_s,
                               const uint64_t *val_s, uint64_t m, uint64_t *sa, uint8_t *keep, uint8_t *head)
{
	const uint32_t FULL = (uint32_t)(32 + (1ll << 31));
	for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x) {
		sa[slot[t]] = val_s[t];
		bool prev = t > 0 && grp_s[t] == grp_s[t - 1] && key_s[t] == key_s[t - 1] && rem_s[t] == FULL && rem_s[t - 1] == FULL;
		bool next = t + 1 < m && grp_s[t] == grp_s[t + 1] && key_s[t] == key_s[t + 1] && rem_s[t] == FULL && rem_s[t + 1] == FULL;
		keep[t] = prev || next;
		head[t] = !prev;
	}
}

// ---- BWT / Occ buckets / SA samples ----
__global__ void k_find_primary(const uint64_t *safull, uint64_t n, unsigned long long *primary)
{
	for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r <= n; r += (uint64_t)gridDim.x * blockDim.x)
		if (safull[r] == 0) *primary = r;
}

__device__ __forceinline__ uint64_t spread32(uint32_t x)
{ // bit i of x -> bit 2i
	uint64_t v = x;
	v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
	v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
	v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
	v = (v | (v << 2)) & 0x3333333333333333ull;
	v = (v | (v << 1)) & 0x5555555555555555ull;
	return v;
}

// one warp per device bucket: the 64 stored-BWT characters x = 64b .. 64b+63 (two per lane)
__global__ void k_bwt_buckets(const uint64_t *W, const uint64_t *safull, uint64_t n, uint64_t primary, uint4 *buckets, uint64_t n_buckets,
                              uint32_t *cntA, uint32_t *cntC, uint32_t *cntG)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_buckets; b += nwarps) {
		uint32_t c[2];
		bool valid[2];
		for (int h = 0; h < 2; ++h) {
			uint64_t x = (b << 6) + 32 * h + lane;
			valid[h] = x < n;
			c[h] = 0;
			if (valid[h]) {
				uint64_t r = x + (x >= primary);     // row of the full matrix ('$' row skipped)
				uint64_t s = safull[r];              // never 0 here
				c[h] = base_at(W, s - 1);
			}
		}
		uint32_t lo0 = __ballot_sync(0xffffffffu, c[0] & 1), hi0 = __ballot_sync(0xffffffffu, c[0] >> 1);
		uint32_t lo1 = __ballot_sync(0xffffffffu, c[1] & 1), hi1 = __ballot_sync(0xffffffffu, c[1] >> 1);
		uint32_t v0 = __ballot_sync(0xffffffffu, valid[0]), v1 = __ballot_sync(0xffffffffu, valid[1]);
		if (lane == 0) {
			uint64_t w0 = spread32(lo0) | (spread32(hi0) << 1), w1 = spread32(lo1) | (spread32(hi1) << 1);
			buckets[2 * b] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
			cntA[b] = __popc(~lo0 & ~hi0 & v0) + __popc(~lo1 & ~hi1 & v1);
			cntC[b] = __popc(lo0 & ~hi0) + __popc(lo1 & ~hi1);
			cntG[b] = __popc(~lo0 & hi0) + __popc(~lo1 & hi1);
		}
	}
}

__global__ void k_bucket_counts(const uint64_t *exA, const uint64_t *exC, const uint64_t *exG, uint4 *buckets, uint64_t n_buckets)
{
	for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < n_buckets; b += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t a = exA[b], c = a + exC[b], g = c + exG[b];   // prefix sums #A, #A+#C, #A+#C+#G (cs_device.cuh)
		buckets[2 * b + 1] = make_uint4((uint32_t)a, (uint32_t)c, (uint32_t)g,
		                                ((uint32_t)(a >> 32) & 0xff) | (((uint32_t)(c >> 32) & 0xff) << 8) | (((uint32_t)(g >> 32) & 0xff) << 16));
	}
}

__global__ void k_widen(const uint32_t *in, uint64_t n, uint64_t *out)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

__global__ void k_sample_sa(const uint64_t *safull, uint64_t n_sa, uint32_t shift, uint64_t *sa)
{
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_sa; i += (uint64_t)gridDim.x * blockDim.x)
		sa[i] = i == 0 ? (uint64_t)-1 : safull[i << shift];
}

template <typename K, typename V>
cudaError_t sort_pairs(void *&tmp, size_t &tmp_bytes, const K *kin, K *kout, const V *vin, V *vout, uint64_t cnt, int end_bit)
{
	size_t need = 0;
	cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, (int64_t)cnt, 0, end_bit);
	if (e != cudaSuccess) return e;
	if (need > tmp_bytes) {
		if (tmp) cudaFree(tmp);
		tmp = nullptr; tmp_bytes = 0;
		if ((e = cudaMalloc(&tmp, need + (need >> 3) + 256)) != cudaSuccess) return e;
		tmp_bytes = need + (need >> 3) + 256;
	}
	return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int64_t)cnt, 0, end_bit);
}

} // namespace

extern "C" cs_index_t *cs_index_build(const uint8_t *fwd, uint64_t l_pac, int device, int sa_intv)
{ return cs_index_build_ex(fwd, l_pac, device, sa_intv, nullptr); }

extern "C" cs_index_t *cs_index_build_ex(const uint8_t *fwd, uint64_t l_pac, int device, int sa_intv, const cs_index_config_t *cfg)
{
	const uint64_t n = 2 * l_pac;
	int n_dev = 0, n_sm = 0, sa_shift = 0;
	uint8_t *d_fwd = nullptr;
	uint64_t *W = nullptr, *safull = nullptr, *d_sa = nullptr;
	unsigned long long *d_small = nullptr;     // [0..15] hist, [16..31] cursors, [32] tie counter, [33] primary, [34] any_short
	unsigned long long h_small[40];
	uint64_t off[17];
	uint64_t *keys_in = nullptr, *keys_out = nullptr, *vals_out = nullptr;
	void *tmp = nullptr; size_t tmp_bytes = 0;
	// tie lists
	uint64_t m = 0, cap = 0;
	uint64_t *t_slot = nullptr, *t_val = nullptr, *t_grp = nullptr; uint8_t *t_head = nullptr;
	uint64_t *r_key = nullptr, *r_key_s = nullptr, *r_grp_s = nullptr, *r_val_s = nullptr, *r_a = nullptr, *r_b = nullptr;
	uint32_t *r_rem = nullptr, *r_rem_s = nullptr, *r_perm = nullptr, *r_perm2 = nullptr, *r_u32a = nullptr, *r_u32b = nullptr;
	uint8_t *r_keep = nullptr;
	unsigned long long *d_nsel = nullptr;
	uint4 *buckets = nullptr;
	uint32_t *cntA = nullptr, *cntC = nullptr, *cntG = nullptr;
	uint64_t *exA = nullptr, *exC = nullptr, *exG = nullptr, *wide = nullptr;
	uint64_t n_buckets = 0, n_words = 0, max_cnt = 0, primary = 0, n_sa = 0;
	uint64_t L2[5];
	cs_index_t *idx = nullptr;
	int grid = 0;

	b_err[0] = 0;
	if (!fwd || l_pac == 0 || n >= (1ull << 37)) { snprintf(b_err, sizeof b_err, "bad reference length %llu", (unsigned long long)l_pac); goto fail; }
	while ((1 << sa_shift) < sa_intv) ++sa_shift;
	if (sa_intv < 1 || (1 << sa_shift) != sa_intv) { snprintf(b_err, sizeof b_err, "sa_intv %d is not a power of two", sa_intv); goto fail; }
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) {
		cudaGetLastError();
		cs_internal_set_error(CS_E_NODEVICE, "no CUDA device visible: compseed_b200 has no CPU path");
		return nullptr;
	}
	BCK(cudaSetDevice(device));
	{
		cudaDeviceProp prop;
		BCK(cudaGetDeviceProperties(&prop, device));
		n_sm = prop.multiProcessorCount;
	}
	grid = n_sm * 8;

	// 1. text
	n_words = (n + 31) / 32 + 2;
	BCK(cudaMalloc(&d_fwd, l_pac));
	BCK(cudaMemcpy(d_fwd, fwd, l_pac, cudaMemcpyHostToDevice));
	BCK(cudaMalloc(&W, n_words * 8));
	k_pack_text<<<grid, 256>>>(d_fwd, l_pac, W, n_words);
	BCK(cudaGetLastError());
	BCK(cudaDeviceSynchronize());
	BCK(cudaFree(d_fwd)); d_fwd = nullptr;

	// 2. MSD partition by the first two bases into safull[1..n] (row 0 is the '$' suffix)
	BCK(cudaMalloc(&safull, (n + 1) * 8));
	BCK(cudaMalloc(&d_small, 40 * 8));
	BCK(cudaMemset(d_small, 0, 40 * 8));
	k_prefix_hist<<<grid, 256>>>(W, n, d_small);
	BCK(cudaGetLastError());
	BCK(cudaMemcpy(h_small, d_small, 16 * 8, cudaMemcpyDeviceToHost));
	off[0] = 0;
	for (int p = 0; p < 16; ++p) { off[p + 1] = off[p] + h_small[p]; max_cnt = std::max<uint64_t>(max_cnt, h_small[p]); }
	for (int p = 0; p < 16; ++p) h_small[16 + p] = off[p];
	BCK(cudaMemcpy(d_small + 16, h_small + 16, 16 * 8, cudaMemcpyHostToDevice));
	k_prefix_scatter<<<grid, 256>>>(W, n, d_small + 16, safull + 1);
	BCK(cudaGetLastError());

	// 3. per-bucket LSD radix sort on the first 32 bases (top 4 bits are constant inside a bucket)
	BCK(cudaMalloc(&keys_in, max_cnt * 8 + 8)); BCK(cudaMalloc(&keys_out, max_cnt * 8 + 8)); BCK(cudaMalloc(&vals_out, max_cnt * 8 + 8));
	for (int p = 0; p < 16; ++p) {
		uint64_t cnt = off[p + 1] - off[p];
		if (cnt == 0) continue;
		uint64_t *slice = safull + 1 + off[p];
		k_keys<<<grid, 256>>>(W, n, slice, cnt, 0, keys_in);
		BCK(cudaGetLastError());
		BCK((sort_pairs<uint64_t, uint64_t>(tmp, tmp_bytes, keys_in, keys_out, slice, vals_out, cnt, 60)));
		BCK(cudaMemcpy(slice, vals_out, cnt * 8, cudaMemcpyDeviceToDevice));
		// ties of this bucket
		BCK(cudaMemset(d_small + 32, 0, 8));
		k_count_ties<<<grid, 256>>>(keys_out, cnt, d_small + 32);
		BCK(cudaGetLastError());
		BCK(cudaMemcpy(h_small + 32, d_small + 32, 8, cudaMemcpyDeviceToHost));
		if (h_small[32]) {
			uint64_t add = h_small[32];
			if (m + add > cap) { // grow the tie lists
				uint64_t ncap = std::max<uint64_t>((m + add) * 2, 1024);
				uint64_t *ns = nullptr, *nv = nullptr; uint8_t *nh = nullptr;
				BCK(cudaMalloc(&ns, ncap * 8)); BCK(cudaMalloc(&nv, ncap * 8)); BCK(cudaMalloc(&nh, ncap));
				if (m) {
					BCK(cudaMemcpy(ns, t_slot, m * 8, cudaMemcpyDeviceToDevice));
					BCK(cudaMemcpy(nv, t_val, m * 8, cudaMemcpyDeviceToDevice));
					BCK(cudaMemcpy(nh, t_head, m, cudaMemcpyDeviceToDevice));
				}
				cudaFree(t_slot); cudaFree(t_val); cudaFree(t_head);
				t_slot = ns; t_val = nv; t_head = nh; cap = ncap;
			}
			BCK(cudaMemset(d_small + 32, 0, 8));
			k_emit_ties<<<grid, 256>>>(keys_out, slice, cnt, 1 + off[p], d_small + 32, t_slot + m, t_val + m, t_head + m);
			BCK(cudaGetLastError());
			m += add;
		}
	}
	cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_out); keys_in = keys_out = vals_out = nullptr;

	// 4. refine ties 32 bases at a time
	if (m) {
		const uint64_t m0 = m;
		BCK(cudaMalloc(&t_grp, m0 * 8)); BCK(cudaMalloc(&r_key, m0 * 8)); BCK(cudaMalloc(&r_key_s, m0 * 8)); BCK(cudaMalloc(&r_grp_s, m0 * 8));
		BCK(cudaMalloc(&r_val_s, m0 * 8)); BCK(cudaMalloc(&r_a, m0 * 8)); BCK(cudaMalloc(&r_b, m0 * 8));
		BCK(cudaMalloc(&r_rem, m0 * 4)); BCK(cudaMalloc(&r_rem_s, m0 * 4)); BCK(cudaMalloc(&r_perm, m0 * 4)); BCK(cudaMalloc(&r_perm2, m0 * 4));
		BCK(cudaMalloc(&r_u32a, m0 * 4)); BCK(cudaMalloc(&r_u32b, m0 * 4)); BCK(cudaMalloc(&r_keep, m0)); BCK(cudaMalloc(&d_nsel, 8));
		int g = (int)std::min<uint64_t>((m0 + 255) / 256, (uint64_t)grid);
		// order the tie list by slot (k_emit_ties appends in arbitrary order): sort (slot -> index), gather
		k_refine_keys<<<g, 256>>>(W, n, t_val, m, 0, r_key, r_rem, r_perm, (unsigned int*)(d_small + 34)); // only for perm = iota
		BCK(cudaGetLastError());
		BCK((sort_pairs<uint64_t, uint32_t>(tmp, tmp_bytes, t_slot, r_a, r_perm, r_perm2, m, 40)));
		k_gather<uint64_t><<<g, 256>>>(t_val, r_perm2, m, r_b);
		k_gather<uint8_t><<<g, 256>>>(t_head, r_perm2, m, r_keep);
		BCK(cudaGetLastError());
		BCK(cudaMemcpy(t_slot, r_a, m * 8, cudaMemcpyDeviceToDevice));
		BCK(cudaMemcpy(t_val, r_b, m * 8, cudaMemcpyDeviceToDevice));
		BCK(cudaMemcpy(t_head, r_keep, m, cudaMemcpyDeviceToDevice));
		for (uint64_t d = 32; m > 0; d += 32) {
			g = (int)std::min<uint64_t>((m + 255) / 256, (uint64_t)grid);
			// group id = slot (+1) of the run head, propagated by an inclusive max-scan
			k_head_slots<<<g, 256>>>(t_slot, t_head, m, r_a);
			BCK(cudaGetLastError());
			{
				size_t need = 0;
				BCK(cub::DeviceScan::InclusiveScan(nullptr, need, r_a, t_grp, MaxOp(), (int64_t)m));
				if (need > tmp_bytes) { if (tmp) cudaFree(tmp); tmp = nullptr; tmp_bytes = 0; BCK(cudaMalloc(&tmp, need + 256)); tmp_bytes = need + 256; }
				BCK(cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, r_a, t_grp, MaxOp(), (int64_t)m));
			}
			BCK(cudaMemset(d_small + 34, 0, 8));
			k_refine_keys<<<g, 256>>>(W, n, t_val, m, d, r_key, r_rem, r_perm, (unsigned int*)(d_small + 34));
			BCK(cudaGetLastError());
			BCK(cudaMemcpy(h_small + 34, d_small + 34, 8, cudaMemcpyDeviceToHost));
			uint32_t *perm = r_perm, *perm_o = r_perm2;
			if (h_small[34] & 0xffffffffull) { // least significant key: remaining length (shorter suffix first)
				BCK((sort_pairs<uint32_t, uint32_t>(tmp, tmp_bytes, r_rem, r_u32a, perm, perm_o, m, 32)));
				std::swap(perm, perm_o);
			}
			k_gather<uint64_t><<<g, 256>>>(r_key, perm, m, r_a);
			BCK(cudaGetLastError());
			BCK((sort_pairs<uint64_t, uint32_t>(tmp, tmp_bytes, r_a, r_b, perm, perm_o, m, 64)));
			std::swap(perm, perm_o);
			k_gather<uint64_t><<<g, 256>>>(t_grp, perm, m, r_a);
			BCK(cudaGetLastError());
			BCK((sort_pairs<uint64_t, uint32_t>(tmp, tmp_bytes, r_a, r_grp_s, perm, perm_o, m, 40)));
			std::swap(perm, perm_o);
			k_gather<uint64_t><<<g, 256>>>(r_key, perm, m, r_key_s);
			k_gather<uint32_t><<<g, 256>>>(r_rem, perm, m, r_rem_s);
			k_gather<uint64_t><<<g, 256>>>(t_val, perm, m, r_val_s);
			k_refine_apply<<<g, 256>>>(t_slot, r_grp_s, r_key_s, r_rem_s, r_val_s, m, safull, r_keep, t_head);
			BCK(cudaGetLastError());
			// compact what is still tied: slot, val, head (all in sorted order; slots stay ascending)
			{
				size_t need = 0;
				BCK(cub::DeviceSelect::Flagged(nullptr, need, t_slot, r_keep, r_a, d_nsel, (int64_t)m));
				if (need > tmp_bytes) { if (tmp) cudaFree(tmp); tmp = nullptr; tmp_bytes = 0; BCK(cudaMalloc(&tmp, need + 256)); tmp_bytes = need + 256; }
				BCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, t_slot, r_keep, r_a, d_nsel, (int64_t)m));
				BCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, r_val_s, r_keep, r_b, d_nsel, (int64_t)m));
				BCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, t_head, r_keep, (uint8_t*)r_u32b, d_nsel, (int64_t)m));
				unsigned long long m_new = 0;
				BCK(cudaMemcpy(&m_new, d_nsel, 8, cudaMemcpyDeviceToHost));
				if (m_new) {
					BCK(cudaMemcpy(t_slot, r_a, m_new * 8, cudaMemcpyDeviceToDevice));
					BCK(cudaMemcpy(t_val, r_b, m_new * 8, cudaMemcpyDeviceToDevice));
					BCK(cudaMemcpy(t_head, r_u32b, m_new, cudaMemcpyDeviceToDevice));
				}
				m = m_new;
			}
			if (d > n + 64) { snprintf(b_err, sizeof b_err, "suffix refinement did not converge"); goto fail; }
		}
	}

	// 5. BWT, Occ buckets, SA samples
	{
		uint64_t m1 = (uint64_t)-1;
		BCK(cudaMemcpy(safull, &m1, 8, cudaMemcpyHostToDevice));   // row 0; never 0, so k_find_primary skips it
		BCK(cudaMemset(d_small + 33, 0, 8));
		k_find_primary<<<grid, 256>>>(safull, n, d_small + 33);
		BCK(cudaGetLastError());
		BCK(cudaMemcpy(h_small + 33, d_small + 33, 8, cudaMemcpyDeviceToHost));
		primary = h_small[33];
		uint64_t nn = n;
		BCK(cudaMemcpy(safull, &nn, 8, cudaMemcpyHostToDevice));   // SA of the '$' row is n: its BWT char is T[n-1]
	}
	n_buckets = (n + 63) / 64 + 1;
	BCK(cudaMalloc(&buckets, n_buckets * 32));
	BCK(cudaMemset(buckets, 0, n_buckets * 32));
	BCK(cudaMalloc(&cntA, n_buckets * 4)); BCK(cudaMalloc(&cntC, n_buckets * 4)); BCK(cudaMalloc(&cntG, n_buckets * 4));
	BCK(cudaMemset(cntA, 0, n_buckets * 4)); BCK(cudaMemset(cntC, 0, n_buckets * 4)); BCK(cudaMemset(cntG, 0, n_buckets * 4));
	k_bwt_buckets<<<grid, 256>>>(W, safull, n, primary, buckets, n_buckets - 1, cntA, cntC, cntG);
	BCK(cudaGetLastError());
	BCK(cudaMalloc(&exA, n_buckets * 8)); BCK(cudaMalloc(&exC, n_buckets * 8)); BCK(cudaMalloc(&exG, n_buckets * 8)); BCK(cudaMalloc(&wide, n_buckets * 8));
	{
		size_t need = 0;
		BCK(cub::DeviceScan::ExclusiveSum(nullptr, need, wide, exA, (int64_t)n_buckets));
		if (need > tmp_bytes) { if (tmp) cudaFree(tmp); tmp = nullptr; tmp_bytes = 0; BCK(cudaMalloc(&tmp, need + 256)); tmp_bytes = need + 256; }
		uint32_t *src[3] = { cntA, cntC, cntG }; uint64_t *dst[3] = { exA, exC, exG };
		for (int c = 0; c < 3; ++c) {
			k_widen<<<grid, 256>>>(src[c], n_buckets, wide);
			BCK(cudaGetLastError());
			BCK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, wide, dst[c], (int64_t)n_buckets));
			BCK(cudaMemcpy(&L2[c + 1], dst[c] + (n_buckets - 1), 8, cudaMemcpyDeviceToHost)); // last bucket is the empty pad: total
		}
	}
	k_bucket_counts<<<grid, 256>>>(exA, exC, exG, buckets, n_buckets);
	BCK(cudaGetLastError());
	// L2: cumulative counts of A, C, G, T
	{
		uint64_t a = L2[1], c = L2[2], g = L2[3];
		L2[0] = 0; L2[1] = a; L2[2] = a + c; L2[3] = a + c + g; L2[4] = n;
	}
	if (sa_intv == 1) {
		uint64_t m1 = (uint64_t)-1;
		BCK(cudaMemcpy(safull, &m1, 8, cudaMemcpyHostToDevice));
		d_sa = safull; safull = nullptr; n_sa = n + 1;
	} else {
		n_sa = (n + sa_intv) / sa_intv;
		BCK(cudaMalloc(&d_sa, n_sa * 8));
		k_sample_sa<<<grid, 256>>>(safull, n_sa, (uint32_t)sa_shift, d_sa);
		BCK(cudaGetLastError());
	}
	BCK(cudaDeviceSynchronize());
	idx = cs_index_adopt(device, buckets, n_buckets, d_sa, n_sa, sa_intv, primary, L2, n, W, cfg);
	buckets = nullptr; d_sa = nullptr;
fail:
	cudaFree(d_fwd); cudaFree(W); cudaFree(safull); cudaFree(d_small); cudaFree(keys_in); cudaFree(keys_out); cudaFree(vals_out);
	cudaFree(tmp); cudaFree(t_slot); cudaFree(t_val); cudaFree(t_grp); cudaFree(t_head);
	cudaFree(r_key); cudaFree(r_key_s); cudaFree(r_grp_s); cudaFree(r_val_s); cudaFree(r_a); cudaFree(r_b);
	cudaFree(r_rem); cudaFree(r_rem_s); cudaFree(r_perm); cudaFree(r_perm2); cudaFree(r_u32a); cudaFree(r_u32b); cudaFree(r_keep); cudaFree(d_nsel);
	cudaFree(cntA); cudaFree(cntC); cudaFree(cntG); cudaFree(exA); cudaFree(exC); cudaFree(exG); cudaFree(wide);
	if (!idx) { cudaFree(buckets); cudaFree(d_sa); cs_internal_set_error(CS_E_CUDA, b_err); }
	return idx;
}
