// Host side of the C-ABI (include/compseed_b200.h): index upload / re-layout, batch contexts with
// pinned double-buffered slots, one CUDA stream per slot, kernel launches and result gathering.
// No CPU compute path exists here: every entry point needs a CUDA device and fails loudly without one.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <vector>
#include <chrono>
#include <thread>
#include <algorithm>
#include <time.h>
#include <cub/device/device_scan.cuh>
#include "cs_kernels.cuh"

#include "cs_internal.h"

// Hardware work queues.  A CUDA context maps its streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8), round robin;
// two streams that share a queue run one after the other.  Measured on the B200 box (scripts/micro/copy_overlap.cu): of 14 streams
// with one kernel each, streams 8..13 start when streams 0..5 end -- and in the slot pipeline the input copy of one slot waited for
// the kernels or the result copy of another (profiles/r02_pipeline_timeline.md).  The variable is read when the context is created,
// so it is set when the library is loaded, unless the host has set it already; a host that creates its CUDA context before loading
// this library sets it itself.
namespace { struct QueueEnv { QueueEnv() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); } } g_queue_env; }

static thread_local char g_err[512] = "";
static thread_local int g_err_code = 0;

extern "C" int cs_last_error_code(void) { return g_err_code; }

int cs_set_err(int code, const char *fmt, ...)
{
	va_list ap;
	g_err_code = code;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}
#define set_err cs_set_err

extern "C" const char *cs_last_error(void) { return g_err; }
void cs_internal_set_error(int code, const char *msg) { set_err(code, "%s", msg); }

extern "C" int cs_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int cs_use_device(int device)
{
	int n = cs_device_count();
	if (n <= 0) return set_err(CS_E_NODEVICE, "no CUDA device visible: compseed_b200 has no CPU path");
	if (device < 0 || device >= n) return set_err(CS_E_ARG, "device %d out of range (%d visible)", device, n);
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess) return set_err(CS_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
	cudaGetLastError();   // an error some earlier call of this thread left behind must not be blamed on the next launch
	return CS_OK;
}
#define use_device cs_use_device

extern "C" void cs_index_config_default(cs_index_config_t *cfg)
{
	if (!cfg) return;
	cfg->kmer_table_depth = -1; cfg->prune_k = -1; cfg->isa_intv = -1; cfg->repeat_lengths = -1;
}

extern "C" void cs_ctx_config_default(cs_ctx_config_t *cfg)
{
	if (!cfg) return;
	cfg->use_fast = -1; cfg->use_r3_fast = -1; cfg->defer_cap = -1; cfg->lit_ctas_per_sm = -1;
	cfg->prefetch_results = 0; cfg->l2_persist_mb = 0; cfg->overlap_streams = 0; cfg->compact_results = 0; cfg->batch_order = -1;
}

static uint64_t kt_offset_host(uint32_t d) { return ((1ull << (2 * d)) - 4) / 3; }   // entries of depths 1 .. d-1 (kt_offset, cs_device.cuh)
static int log2_exact(uint64_t v) { int s = 0; while ((1ull << s) < v) ++s; return (1ull << s) == v ? s : -1; }

// Unique-match fast path (cs_device.cuh): 2-bit text in read bit order + sampled inverse SA.  Only with a
// dense SA.  cfg.isa_intv sets the ISA sampling (power of two, default 2; 0 disables the fast path).
static int cs_internal_build_text(cs_index *idx, const uint64_t *W)
{
	int intv = idx->cfg.isa_intv < 0 ? 2 : idx->cfg.isa_intv, shift = 0;   // default 2: 4 more bytes per row than 4, 7 % off a cfg2 step (fewer LF steps after every inverse-SA lookup)
	const uint64_t n = idx->d.seq_len;
	idx->d.text = nullptr; idx->d.isa = nullptr; idx->d.isa_shift = 0; idx->d_text = nullptr; idx->d_isa = nullptr;
	idx->d.rep = nullptr; idx->d_rep = nullptr;
	if (intv <= 0 || idx->d.sa_mask != 0 || !W) return CS_OK;
	while ((1 << shift) < intv) ++shift;
	{
		const uint64_t n_words = (n + 31) / 32 + 2, n_isa = (n >> shift) + 2;
		int grid = idx->n_sm * 8;
		CK(cudaMalloc(&idx->d_text, n_words * 8));
		k_text_lsb<<<grid, 256>>>(W, n_words, idx->d_text);
		CK(cudaGetLastError());
		CK(cudaMalloc(&idx->d_isa, n_isa * 8));
		CK(cudaMemset(idx->d_isa, 0, n_isa * 8));
		k_isa_sample<<<grid, 256>>>(idx->d, idx->d_isa, (uint32_t)shift);
		CK(cudaGetLastError());
		CK(cudaDeviceSynchronize());
		idx->d.text = idx->d_text; idx->d.isa = idx->d_isa; idx->d.isa_shift = (uint32_t)shift;
		idx->bytes += n_words * 8 + n_isa * 8;
		if (idx->cfg.repeat_lengths != 0) { // repeat lengths: the suffix array's neighbours compared through the text (cs_device.cuh)
			const uint64_t n_rep = ((n + 63) & ~63ull) + 64;   // k_seed_fast reads whole 8-byte words around a position
			CK(cudaMalloc(&idx->d_rep, n_rep));
			CK(cudaMemset(idx->d_rep, 0, n_rep));
			k_rep_build<<<grid, 256>>>(idx->d, idx->d_rep);
			CK(cudaGetLastError());
			CK(cudaDeviceSynchronize());
			idx->d.rep = idx->d_rep;
			idx->bytes += n_rep;
		}
	}
	return CS_OK;
fail:
	cudaFree(idx->d_text); cudaFree(idx->d_isa); cudaFree(idx->d_rep); idx->d_text = idx->d_isa = nullptr; idx->d_rep = nullptr;
	idx->d.text = nullptr; idx->d.isa = nullptr; idx->d.rep = nullptr;
	return CS_E_CUDA;
}

// Occurrence filter (cs_device.cuh, "Occurrence filter" at k_seed): 2-bit saturating counts of all K-mers of the
// indexed text, K = ceil(log4(seq_len)) + 2 capped at 19 (17 GB at K = 18, 69 GB at K = 19): long enough
// that a random K-mer is almost always absent, short enough to stay below the default min_seed_len.
// W is the 2-bit text if the caller has it (the on-device builder); otherwise it is rebuilt from the
// BWT and the SA.  cfg.prune_k overrides (0 disables).
int cs_internal_build_filter(cs_index *idx, const uint64_t *W)
{
	unsigned long long *own = nullptr;
	int K = 2;
	const uint64_t n = idx->d.seq_len;
	const int grid = idx->n_sm * 8;
	while (K < 19 && (1ull << (2 * (K - 2))) < n) ++K;      // K - 2 >= log4(n)
	if (K < 8) K = 8;
	if (idx->cfg.prune_k >= 0) K = idx->cfg.prune_k;
	if (K > 0 && K < 4) K = 4;
	if (K > 19) K = 19;
	if (K < 0 || n < (uint64_t)K) K = 0;
	idx->d.pt = nullptr; idx->d.pt_k = 0; idx->d_pt = nullptr;
	idx->d.text = nullptr; idx->d.isa = nullptr; idx->d.isa_shift = 0; idx->d_text = nullptr; idx->d_isa = nullptr;
	if (!W) { // rebuild the 2-bit text from the BWT and the SA
		const uint64_t n_words = (n + 31) / 32 + 2;
		CK(cudaMalloc(&own, n_words * 8));
		CK(cudaMemset(own, 0, n_words * 8));
		k_text_from_index<<<grid, 256>>>(idx->d, own);
		CK(cudaGetLastError());
		W = reinterpret_cast<const uint64_t*>(own);
	}
	if (K > 0) {
		const uint64_t words = (1ull << (2 * K)) / 16;
		CK(cudaMalloc(&idx->d_pt, words * 4));
		CK(cudaMemset(idx->d_pt, 0, words * 4));
		k_pt_count<<<grid, 256>>>(W, n, (uint32_t)K, idx->d_pt);
		CK(cudaGetLastError());
		CK(cudaDeviceSynchronize());
		idx->d.pt = idx->d_pt; idx->d.pt_k = (uint32_t)K;
		idx->bytes += words * 4;
	}
	if (cs_internal_build_text(idx, W) != CS_OK) goto fail;
	if (own) cudaFree(own);
	return CS_OK;
fail:
	if (own) cudaFree(own);
	if (idx->d_pt) { cudaFree(idx->d_pt); idx->d_pt = nullptr; }
	return CS_E_CUDA;
}

// Top-of-search table (cs_device.cuh): depth chosen so that the deepest level still has ~64 rows per
// entry (deeper levels cost HBM without saving sector reads: k and l already share a bucket there).
// cfg.kmer_table_depth overrides (0 disables).
static int build_kmer_table(cs_index *idx)
{
	int depth = 0;
	while (depth < 13 && (idx->d.seq_len >> (2 * (depth + 1))) >= 64) ++depth;
	if (idx->cfg.kmer_table_depth >= 0) depth = idx->cfg.kmer_table_depth;
	if (depth < 0) depth = 0;
	if (depth > 15) depth = 15;
	idx->d.kt = nullptr; idx->d.kt_depth = 0; idx->d_kt = nullptr;
	if (depth == 0) return CS_OK;
	uint64_t entries = ((1ull << (2 * (depth + 1))) - 4) / 3;
	CK(cudaMalloc(&idx->d_kt, entries * sizeof(uint4)));
	for (int d = 1; d <= depth; ++d) {
		uint64_t n_parent = 1ull << (2 * (d - 1));
		int grid = (int)std::min<uint64_t>((n_parent + 255) / 256, (uint64_t)idx->n_sm * 16);
		k_kt_build<<<grid, 256>>>(idx->d, idx->d_kt, (uint32_t)d);
		CK(cudaGetLastError());
	}
	CK(cudaDeviceSynchronize());
	idx->d.kt = idx->d_kt; idx->d.kt_depth = (uint32_t)depth;
	idx->bytes += entries * sizeof(uint4);
	return CS_OK;
fail:
	if (idx->d_kt) { cudaFree(idx->d_kt); idx->d_kt = nullptr; }
	return CS_E_CUDA;
}

// replaces the device SA by one sampled every new_intv rows
static int resample_sa(cs_index *idx, int new_intv)
{
	uint64_t *d_new = nullptr;
	int sh = log2_exact((uint64_t)new_intv);
	if (sh < 0) return set_err(CS_E_ARG, "SA interval %d is not a power of two", new_intv);
	uint64_t n_new = (idx->d.seq_len + new_intv) / new_intv;
	CK(cudaMalloc(&d_new, n_new * 8));
	k_resample_sa<<<idx->n_sm * 8, 256>>>(idx->d, d_new, n_new, (uint32_t)sh);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	CK(cudaFree(idx->d_sa));
	idx->bytes += n_new * 8; idx->bytes -= idx->d.n_sa * 8;
	idx->d_sa = d_new; idx->d.sa = d_new; idx->d.n_sa = n_new;
	idx->d.sa_mask = (uint32_t)new_intv - 1; idx->d.sa_shift = (uint32_t)sh; idx->sa_intv = new_intv;
	return CS_OK;
fail:
	if (d_new) cudaFree(d_new);
	return CS_E_CUDA;
}

extern "C" cs_index_t *cs_index_upload(const cs_bwt_view_t *v, int device, int dense_sa_intv)
{ return cs_index_upload_ex(v, device, dense_sa_intv, nullptr); }

extern "C" cs_index_t *cs_index_upload_ex(const cs_bwt_view_t *v, int device, int dense_sa_intv, const cs_index_config_t *cfg)
{
	cs_index *idx = nullptr;
	uint32_t *d_src = nullptr;
	if (!v || !v->bwt || !v->sa) { set_err(CS_E_ARG, "null bwt view"); return nullptr; }
	if (v->seq_len == 0 || v->seq_len >= (1ull << 37)) { set_err(CS_E_ARG, "seq_len %llu outside (0, 2^37)", (unsigned long long)v->seq_len); return nullptr; }
	if (log2_exact((uint64_t)v->sa_intv) < 0) { set_err(CS_E_ARG, "sa_intv %d is not a power of two", v->sa_intv); return nullptr; }
	if (use_device(device) != CS_OK) return nullptr;
	if (v->n_sa != (v->seq_len + (uint64_t)v->sa_intv) / (uint64_t)v->sa_intv) { set_err(CS_E_ARG, "n_sa %llu does not match seq_len / sa_intv (bwt.c:435)", (unsigned long long)v->n_sa); return nullptr; }
	if (v->bwt_size < ((v->seq_len + 15) >> 4)) { set_err(CS_E_ARG, "bwt_size %llu too small for seq_len %llu", (unsigned long long)v->bwt_size, (unsigned long long)v->seq_len); return nullptr; }
	idx = (cs_index*)calloc(1, sizeof(cs_index));
	if (!idx) { set_err(CS_E_ARG, "out of host memory"); return nullptr; }
	idx->device = device;
	if (cfg) idx->cfg = *cfg; else cs_index_config_default(&idx->cfg);
	{
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		idx->n_sm = prop.multiProcessorCount;
	}
	idx->bwt_size_ref = v->bwt_size;
	idx->d.primary = v->primary; idx->d.seq_len = v->seq_len;
	for (int i = 0; i < 5; ++i) idx->d.L2[i] = v->L2[i];
	idx->d.n_buckets = (v->seq_len + 63) / 64 + 1; // one zero pad bucket: row `primary` may index one past the end (bwt.c:55-58)
	CK(cudaMalloc(&idx->d_buckets, idx->d.n_buckets * 32));
	CK(cudaMemset(idx->d_buckets, 0, idx->d.n_buckets * 32));
	CK(cudaMalloc(&d_src, v->bwt_size * 4));
	CK(cudaMemcpy(d_src, v->bwt, v->bwt_size * 4, cudaMemcpyHostToDevice));
	k_relayout<<<idx->n_sm * 8, 256>>>(d_src, v->bwt_size, v->seq_len, idx->d_buckets, idx->d.n_buckets - 1);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	CK(cudaFree(d_src)); d_src = nullptr;
	idx->d.buckets = idx->d_buckets;
	CK(cudaMalloc(&idx->d_sa, v->n_sa * 8));
	CK(cudaMemcpy(idx->d_sa, v->sa, v->n_sa * 8, cudaMemcpyHostToDevice));
	{
		uint64_t m1 = (uint64_t)-1;
		CK(cudaMemcpy(idx->d_sa, &m1, 8, cudaMemcpyHostToDevice)); // sa[0] = -1 (bwt.c:83,437)
	}
	idx->d.sa = idx->d_sa; idx->d.n_sa = v->n_sa;
	idx->sa_intv = v->sa_intv;
	idx->d.sa_mask = (uint32_t)v->sa_intv - 1; idx->d.sa_shift = (uint32_t)log2_exact((uint64_t)v->sa_intv);
	idx->bytes = idx->d.n_buckets * 32 + v->n_sa * 8;
	if (dense_sa_intv > 0 && dense_sa_intv < v->sa_intv)
		if (resample_sa(idx, dense_sa_intv) != CS_OK) goto fail;
	if (build_kmer_table(idx) != CS_OK) goto fail;
	if (cs_internal_build_filter(idx, nullptr) != CS_OK) goto fail;
	return idx;
fail:
	if (d_src) cudaFree(d_src);
	if (idx) { if (idx->d_buckets) cudaFree(idx->d_buckets); if (idx->d_sa) cudaFree(idx->d_sa); if (idx->d_kt) cudaFree(idx->d_kt); if (idx->d_pt) cudaFree(idx->d_pt); free(idx); }
	return nullptr;
}

extern "C" cs_index_t *cs_index_load(const char *prefix, int device, int dense_sa_intv)
{ return cs_index_load_ex(prefix, device, dense_sa_intv, nullptr); }

static int64_t file_size(FILE *fp)
{
	if (fseek(fp, 0, SEEK_END) != 0) return -1;
	const int64_t n = (int64_t)ftell(fp);
	if (fseek(fp, 0, SEEK_SET) != 0) return -1;
	return n;
}

extern "C" cs_index_t *cs_index_load_ex(const char *prefix, int device, int dense_sa_intv, const cs_index_config_t *cfg)
{ // file format: bwt.c:385-407 (dump), bwt.c:421-462 (restore).  Corrupt or truncated files give CS_E_IO, never a crash.
	char fn[4096];
	cs_bwt_view_t v;
	memset(&v, 0, sizeof v);
	uint32_t *bwt = nullptr; uint64_t *sa = nullptr;
	cs_index_t *idx = nullptr;
	FILE *fp = nullptr;
	uint64_t hdr[7];
	int64_t fsz;
	if (!prefix) { set_err(CS_E_ARG, "null prefix"); return nullptr; }
	if (cs_device_count() <= 0) { set_err(CS_E_NODEVICE, "no CUDA device visible: compseed_b200 has no CPU path"); return nullptr; }
	snprintf(fn, sizeof fn, "%s.bwt", prefix);
	if ((fp = fopen(fn, "rb")) == nullptr) { set_err(CS_E_IO, "cannot open %s", fn); return nullptr; }
	fsz = file_size(fp);
	if (fsz < 40 + 4 || ((fsz - 40) & 3)) { set_err(CS_E_IO, "%s: %lld bytes is not a .bwt file (40-byte header + 32-bit words)", fn, (long long)fsz); goto done; }
	v.bwt_size = (uint64_t)(fsz - 40) >> 2;
	bwt = (uint32_t*)malloc(v.bwt_size * 4);
	if (!bwt) { set_err(CS_E_IO, "out of host memory reading %s", fn); goto done; }
	if (fread(&v.primary, 8, 1, fp) != 1 || fread(v.L2 + 1, 8, 4, fp) != 4 || fread(bwt, 4, v.bwt_size, fp) != v.bwt_size) {
		set_err(CS_E_IO, "short read on %s", fn); goto done;
	}
	fclose(fp); fp = nullptr;
	v.seq_len = v.L2[4];
	if (v.seq_len == 0 || v.primary > v.seq_len || v.L2[1] > v.L2[2] || v.L2[2] > v.L2[3] || v.L2[3] > v.L2[4] ||
	    v.bwt_size != ((v.seq_len + 15) >> 4) + ((v.seq_len + 127) / 128 + 1) * 8) {
		set_err(CS_E_IO, "%s: header and size disagree (seq_len %llu, %llu words)", fn, (unsigned long long)v.seq_len, (unsigned long long)v.bwt_size); goto done;
	}
	snprintf(fn, sizeof fn, "%s.sa", prefix);
	if ((fp = fopen(fn, "rb")) == nullptr) { set_err(CS_E_IO, "cannot open %s", fn); goto done; }
	fsz = file_size(fp);
	if (fsz < 56 || fread(hdr, 8, 7, fp) != 7) { set_err(CS_E_IO, "short read on %s", fn); goto done; }
	if (hdr[0] != v.primary || hdr[6] != v.seq_len) { set_err(CS_E_IO, "SA-BWT inconsistency in %s (bwt.c:429,433)", fn); goto done; }
	if (hdr[5] == 0 || hdr[5] > (1u << 20) || log2_exact(hdr[5]) < 0) { set_err(CS_E_IO, "%s: sa_intv %llu is not a power of two", fn, (unsigned long long)hdr[5]); goto done; }
	v.sa_intv = (int32_t)hdr[5];
	v.n_sa = (v.seq_len + v.sa_intv) / v.sa_intv;
	if ((uint64_t)(fsz - 56) != (v.n_sa - 1) * 8) { set_err(CS_E_IO, "%s: %lld bytes, expected %llu (bwt.c:435)", fn, (long long)fsz, (unsigned long long)(56 + (v.n_sa - 1) * 8)); goto done; }
	sa = (uint64_t*)malloc(v.n_sa * 8);
	if (!sa) { set_err(CS_E_IO, "out of host memory reading %s", fn); goto done; }
	sa[0] = (uint64_t)-1;
	if (fread(sa + 1, 8, v.n_sa - 1, fp) != v.n_sa - 1) { set_err(CS_E_IO, "short read on %s", fn); goto done; }
	fclose(fp); fp = nullptr;
	v.bwt = bwt; v.sa = sa;
	idx = cs_index_upload_ex(&v, device, dense_sa_intv, cfg);
done:
	if (fp) fclose(fp);
	free(bwt); free(sa);
	return idx;
}

extern "C" int cs_index_write(const cs_index_t *idx, const char *prefix, int sa_intv)
{ // bwt_dump_bwt / bwt_dump_sa, FM_index/bwt.c:385-407
	char fn[4096];
	cs_bwt_view_t v;
	uint32_t *bwt = nullptr; uint64_t *sa = nullptr;
	FILE *fp = nullptr;
	int rc;
	if (!idx || !prefix) return set_err(CS_E_ARG, "null argument");
	if ((rc = cs_index_download(idx, &v, nullptr, nullptr, sa_intv)) != CS_OK) return rc;
	bwt = (uint32_t*)malloc(v.bwt_size * 4); sa = (uint64_t*)malloc(v.n_sa * 8);
	if (!bwt || !sa) { free(bwt); free(sa); return set_err(CS_E_IO, "out of host memory"); }
	if ((rc = cs_index_download(idx, &v, bwt, sa, sa_intv)) != CS_OK) { free(bwt); free(sa); return rc; }
	rc = CS_E_IO;
	snprintf(fn, sizeof fn, "%s.bwt", prefix);
	if ((fp = fopen(fn, "wb")) == nullptr) { set_err(CS_E_IO, "cannot create %s", fn); goto done; }
	if (fwrite(&v.primary, 8, 1, fp) != 1 || fwrite(v.L2 + 1, 8, 4, fp) != 4 || fwrite(bwt, 4, v.bwt_size, fp) != v.bwt_size) { set_err(CS_E_IO, "short write on %s", fn); goto done; }
	if (fclose(fp) != 0) { fp = nullptr; set_err(CS_E_IO, "close failed on %s", fn); goto done; }
	fp = nullptr;
	snprintf(fn, sizeof fn, "%s.sa", prefix);
	if ((fp = fopen(fn, "wb")) == nullptr) { set_err(CS_E_IO, "cannot create %s", fn); goto done; }
	{
		const uint64_t intv = (uint64_t)sa_intv;
		if (fwrite(&v.primary, 8, 1, fp) != 1 || fwrite(v.L2 + 1, 8, 4, fp) != 4 || fwrite(&intv, 8, 1, fp) != 1 || fwrite(&v.seq_len, 8, 1, fp) != 1 ||
		    fwrite(sa + 1, 8, v.n_sa - 1, fp) != v.n_sa - 1) { set_err(CS_E_IO, "short write on %s", fn); goto done; }
	}
	if (fclose(fp) != 0) { fp = nullptr; set_err(CS_E_IO, "close failed on %s", fn); goto done; }
	fp = nullptr;
	rc = CS_OK;
done:
	if (fp) fclose(fp);
	free(bwt); free(sa);
	return rc;
}

extern "C" int cs_index_info(const cs_index_t *idx, cs_bwt_view_t *v, uint64_t *device_bytes)
{
	if (!idx) return set_err(CS_E_ARG, "null index");
	if (v) {
		memset(v, 0, sizeof *v);
		v->primary = idx->d.primary; v->seq_len = idx->d.seq_len;
		for (int i = 0; i < 5; ++i) v->L2[i] = idx->d.L2[i];
		v->bwt_size = idx->bwt_size_ref; v->sa_intv = idx->sa_intv; v->n_sa = idx->d.n_sa;
	}
	if (device_bytes) *device_bytes = idx->bytes;
	return CS_OK;
}

extern "C" int cs_index_download(const cs_index_t *idx, cs_bwt_view_t *v, uint32_t *bwt, uint64_t *sa, int out_sa_intv)
{
	uint32_t *d_ref = nullptr; uint64_t *d_sa = nullptr;
	if (!idx || !v) return set_err(CS_E_ARG, "null argument");
	int sh = log2_exact((uint64_t)out_sa_intv);
	if (sh < 0) return set_err(CS_E_ARG, "out_sa_intv %d is not a power of two", out_sa_intv);
	{ const int rc_ = use_device(idx->device); if (rc_ != CS_OK) return rc_; }
	cs_index_info(idx, v, nullptr);
	v->sa_intv = out_sa_intv;
	v->n_sa = (idx->d.seq_len + out_sa_intv) / out_sa_intv;
	if (!bwt && !sa) return CS_OK;
	if (bwt) {
		CK(cudaMalloc(&d_ref, idx->bwt_size_ref * 4));
		CK(cudaMemset(d_ref, 0, idx->bwt_size_ref * 4));
		k_unlayout<<<idx->n_sm * 8, 256>>>(idx->d_buckets, idx->d.seq_len, d_ref, idx->bwt_size_ref);
		CK(cudaGetLastError());
		CK(cudaMemcpy(bwt, d_ref, idx->bwt_size_ref * 4, cudaMemcpyDeviceToHost));
		CK(cudaFree(d_ref)); d_ref = nullptr;
	}
	if (sa) {
		if (out_sa_intv == idx->sa_intv) CK(cudaMemcpy(sa, idx->d_sa, v->n_sa * 8, cudaMemcpyDeviceToHost));
		else {
			CK(cudaMalloc(&d_sa, v->n_sa * 8));
			k_resample_sa<<<idx->n_sm * 8, 256>>>(idx->d, d_sa, v->n_sa, (uint32_t)sh);
			CK(cudaGetLastError());
			CK(cudaMemcpy(sa, d_sa, v->n_sa * 8, cudaMemcpyDeviceToHost));
			CK(cudaFree(d_sa)); d_sa = nullptr;
		}
	}
	return CS_OK;
fail:
	if (d_ref) cudaFree(d_ref);
	if (d_sa) cudaFree(d_sa);
	return CS_E_CUDA;
}

extern "C" void cs_index_free(cs_index_t *idx)
{
	if (!idx) return;
	cudaSetDevice(idx->device);
	cudaFree(idx->d_buckets); cudaFree(idx->d_sa); cudaFree(idx->d_kt); cudaFree(idx->d_pt); cudaFree(idx->d_text); cudaFree(idx->d_isa); cudaFree(idx->d_rep);
	free(idx);
}

// internal: wrap device arrays produced by the on-device builder (cs_index_build.cu)
cs_index_t *cs_index_adopt(int device, uint4 *d_buckets, uint64_t n_buckets, uint64_t *d_sa, uint64_t n_sa, int sa_intv,
                           uint64_t primary, const uint64_t L2[5], uint64_t seq_len, const uint64_t *W, const cs_index_config_t *cfg)
{
	cs_index *idx = (cs_index*)calloc(1, sizeof(cs_index));
	cudaDeviceProp prop;
	if (cfg) idx->cfg = *cfg; else cs_index_config_default(&idx->cfg);
	cudaGetDeviceProperties(&prop, device);
	idx->device = device; idx->n_sm = prop.multiProcessorCount;
	idx->d_buckets = d_buckets; idx->d_sa = d_sa;
	idx->d.buckets = d_buckets; idx->d.n_buckets = n_buckets;
	idx->d.primary = primary; idx->d.seq_len = seq_len;
	for (int i = 0; i < 5; ++i) idx->d.L2[i] = L2[i];
	idx->d.sa = d_sa; idx->d.n_sa = n_sa; idx->sa_intv = sa_intv;
	idx->d.sa_mask = (uint32_t)sa_intv - 1; idx->d.sa_shift = (uint32_t)log2_exact((uint64_t)sa_intv);
	idx->bwt_size_ref = ((seq_len + 15) >> 4) + ((seq_len + 127) / 128 + 1) * 8;
	idx->bytes = n_buckets * 32 + n_sa * 8;
	if (build_kmer_table(idx) != CS_OK || cs_internal_build_filter(idx, W) != CS_OK) {
		cudaFree(idx->d_kt); cudaFree(d_buckets); cudaFree(d_sa); free(idx);
		return nullptr;
	}
	return idx;
}

// ---------------------------------------------------------------------------------------------
// probes
// ---------------------------------------------------------------------------------------------
extern "C" int cs_occ4(const cs_index_t *idx, uint32_t n, const uint64_t *k, uint64_t *cnt)
{
	uint64_t *d_k = nullptr, *d_c = nullptr;
	if (!idx || !k || !cnt) return set_err(CS_E_ARG, "null argument");
	if (n == 0) return CS_OK;
	{ const int rc_ = use_device(idx->device); if (rc_ != CS_OK) return rc_; }
	CK(cudaMalloc(&d_k, (size_t)n * 8)); CK(cudaMalloc(&d_c, (size_t)n * 32));
	CK(cudaMemcpy(d_k, k, (size_t)n * 8, cudaMemcpyHostToDevice));
	k_probe_occ4<<<(n + 255) / 256, 256>>>(idx->d, n, d_k, d_c);
	CK(cudaGetLastError());
	CK(cudaMemcpy(cnt, d_c, (size_t)n * 32, cudaMemcpyDeviceToHost));
	cudaFree(d_k); cudaFree(d_c);
	return CS_OK;
fail:
	cudaFree(d_k); cudaFree(d_c);
	return CS_E_CUDA;
}

extern "C" int cs_extend(const cs_index_t *idx, uint32_t n, const uint64_t *ik, const int32_t *is_back, uint64_t *ok)
{
	uint64_t *d_ik = nullptr, *d_ok = nullptr; int32_t *d_b = nullptr;
	if (!idx || !ik || !is_back || !ok) return set_err(CS_E_ARG, "null argument");
	if (n == 0) return CS_OK;
	{ const int rc_ = use_device(idx->device); if (rc_ != CS_OK) return rc_; }
	CK(cudaMalloc(&d_ik, (size_t)n * 24)); CK(cudaMalloc(&d_ok, (size_t)n * 96)); CK(cudaMalloc(&d_b, (size_t)n * 4));
	CK(cudaMemcpy(d_ik, ik, (size_t)n * 24, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_b, is_back, (size_t)n * 4, cudaMemcpyHostToDevice));
	k_probe_extend<<<(n + 255) / 256, 256>>>(idx->d, n, d_ik, d_b, d_ok);
	CK(cudaGetLastError());
	CK(cudaMemcpy(ok, d_ok, (size_t)n * 96, cudaMemcpyDeviceToHost));
	cudaFree(d_ik); cudaFree(d_ok); cudaFree(d_b);
	return CS_OK;
fail:
	cudaFree(d_ik); cudaFree(d_ok); cudaFree(d_b);
	return CS_E_CUDA;
}

extern "C" int cs_sa(const cs_index_t *idx, uint32_t n, const uint64_t *k, uint64_t *out)
{
	uint64_t *d_k = nullptr; unsigned long long *d_w = nullptr;
	if (!idx || !k || !out) return set_err(CS_E_ARG, "null argument");
	if (n == 0) return CS_OK;
	{ const int rc_ = use_device(idx->device); if (rc_ != CS_OK) return rc_; }
	CK(cudaMalloc(&d_k, (size_t)n * 8)); CK(cudaMalloc(&d_w, 24));
	CK(cudaMemset(d_w, 0, 24));
	CK(cudaMemcpy(d_w + 2, &n, 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_k, k, (size_t)n * 8, cudaMemcpyHostToDevice));
	k_sa_resolve<<<idx->n_sm * 4, 256>>>(idx->d, reinterpret_cast<const uint32_t*>(d_w + 2), n, d_k, d_w, d_w + 1);
	CK(cudaGetLastError());
	CK(cudaMemcpy(out, d_k, (size_t)n * 8, cudaMemcpyDeviceToHost));
	cudaFree(d_k); cudaFree(d_w);
	return CS_OK;
fail:
	cudaFree(d_k); cudaFree(d_w);
	return CS_E_CUDA;
}

// ---------------------------------------------------------------------------------------------
// batch contexts
// ---------------------------------------------------------------------------------------------
struct Ctrl { // zeroed before every run; copied back after it
	uint32_t next_read[4];   // work counters: [0] k_seed, [1] k_seed_r3, [2] k_seed_fast, [3] k_seed_walk
	uint32_t n_defer;        // calls handed on by k_seed_fast (and k_seed_walk)
	uint32_t n_lit;          // of those, for k_seed
	uint32_t n_defer_fast;   // n_defer when k_seed_fast ended
	uint32_t n_walk;         // walk tasks among them
	unsigned long long pool_used;
	unsigned long long counters[4 + 32];   // [4..19] k_seed, [20..35] k_seed_fast: event counters of a -DCS_STATS diagnostics build, else 0
	unsigned long long sa_work, lf_steps;
	unsigned long long req[6];             // executed memory requests per kernel (SeedArgs::req)
	unsigned long long tot12, tot3, tot_seeds;   // 64-bit batch totals: mems of passes 1-2, of pass 3, seeds
	int error, pad1;
	uint32_t walk_hist[64], walk_cursor[64];   // counting sort of the walk tasks by expected length
	uint32_t n_mems, n_seeds;
	unsigned long long chain_work;         // work counter of k_chain_build
	uint32_t n_chains, n_cseeds, n_nodes, pad2;
};

struct Slot {
	cudaStream_t stream;
	// Result copies (and the control block) go through a stream of their own.  Measured (scripts/micro/pipe_mimic.cu,
	// profiles/r02_pipeline_timeline.md): a stream whose first copy was host-to-device keeps ALL its copies on that copy engine, so
	// with one stream per slot every input and result copy of every slot queued up on one engine, in order of submission: the input
	// copy of a resubmitted slot waited for the result copies of all other slots, and no kernel ran meanwhile.
	cudaStream_t stream_out;
	cudaStream_t stream2;  // the third-pass kernel runs here, next to k_seed_walk / k_seed (it depends on k_seed_fast only)
	cudaEvent_t ev[8];   // slot start, seed start, seed end, collect end, sa end, k_seed end, k_seed_fast end, k_seed_walk end
	cudaEvent_t ev_pack;   // k_pack_reads done
	cudaEvent_t ev_kend;   // last kernel of the batch done (timed twin of ev_kdone)
	cudaEvent_t ev_fork, ev_join, ev_r3[2];   // fork / join of stream2; start and end of the third-pass kernel on it
	cudaEvent_t ev_done;
	cudaEvent_t ev_copy;   // result copies of the batch enqueued (cs_i_fetch_*_into): diagnostics timeline
	cudaEvent_t ev_kdone;  // kernels and the control block copy of the batch in flight are complete
	bool want_fetch;       // submitted through cs_seed_batch_submit: the host wants the results (see prefetch_ready)
	// pinned host
	uint8_t *h_bases; uint32_t *h_off;
	uint32_t *h_mem_off, *h_seed_off; cs_mem_t *h_mems; int64_t *h_rbeg;
	cs_cmem_t *h_cmems; uint32_t *h_rlo; uint8_t *h_rhi;   // compact wire format (pinned, allocated on first use)
	cs_cmem_t *d_cmems; uint32_t *d_rlo; uint8_t *d_rhi;   // ... on the device (cfg.compact_results)
	bool fetched_compact;
	// chaining (cs_ctx_set_chaining): scratch and results on the device, pinned results on first use
	uint32_t *d_s_next, *d_s_qb_len, *d_order, *d_klist, *d_node_cnt, *d_node_off, *d_nodes, *d_n_chain, *d_n_cseed, *d_l_rep, *d_chain_off, *d_cseed_off;
	ChainTmp *d_chain_tmp; cs_chain_t *d_chains; uint32_t *d_cs_lo; uint8_t *d_cs_hi; uint16_t *d_cs_qbeg, *d_cs_len;
	uint32_t *h_chain_off, *h_cseed_off; cs_chain_t *h_chains; uint32_t *h_cs_lo; uint8_t *h_cs_hi; uint16_t *h_cs_qbeg, *h_cs_len;
	bool chained;          // the batch in the slot was chained
	Ctrl *h_ctrl;
	// device
	uint8_t *d_bases; uint32_t *d_off;
	uint64_t *d_packed; uint32_t *d_nmask;   // 2-bit packed reads + ambiguity mask (k_pack_reads)
	uint4 *d_defer_q;                         // calls the fast kernel hands to the literal kernel (SeedArgs::defer_q)
	uint32_t *d_read_last_q, *d_x_n, *d_defer_bits, *d_lit_q; uint64_t *d_x_off; uint4 *d_defer_lx, *d_thread_lx; uint32_t *d_walk_order;
	cs_mem_t *d_stage;                        // collect: a read's sources gathered before the sort
	bool used_fast;
	bool packed_input;                        // the batch came through cs_seed_batch_submit_packed: d_packed / d_nmask are the input
	uint32_t off_bias;                        // ... as a slice of a larger packed set (SeedArgs::off_bias); 0 otherwise
	uint64_t *h_packed; uint32_t *h_nmask;    // pinned staging for it (allocated on first use)
	Ctrl *d_ctrl;
	cs_mem_t *d_thread_mems; uint4 *d_spill;
	cs_mem_t *d_pool, *d_mems;
	uint64_t *d_read_pool_off;
	uint32_t *d_read_n_mems, *d_mem_off, *d_read_n_seeds, *d_seed_off;
	uint64_t *d_rows;
	cs_mem_t *d_r3_mems; uint64_t r3_cap;   // third-pass seeds, read r at off[r]/(k+1) + r
	uint32_t *d_r3_n_mems, *d_tot_n_mems;
	void *d_scan_tmp; size_t scan_tmp_bytes;
	// state
	int state;           // 0 idle, 1 staged, 2 running, 3 done (results on device)
	uint32_t n_reads;
	cs_seed_opt_t opt;
};

struct cs_ctx {
	const cs_index *idx;
	int device;           // (the ctx may outlive a careless caller's index handle: never read idx->device when freeing)
	uint32_t max_reads, max_read_len;
	uint64_t max_bases, max_mems, max_seeds;
	int n_slots;
	int grid;             // CTAs of k_seed
	int grid_fast;        // CTAs of k_seed_fast (0: the index or the read length does not allow the fast kernel)
	int grid_r3;          // CTAs of k_seed_r3
	uint32_t mem_cap, spill_cap, defer_cap;
	cs_ctx_config_t cfg;
	uint64_t n_launch;    // kernels launched so far (counted at the launch sites)
	uint64_t need_mems[16], need_seeds[16];   // per slot: what the last overflowing batch would have needed
	// chaining (cs_ctx_set_chaining)
	bool chaining;
	cs_chain_opt_t copt;
	int64_t l_pac; int32_t n_seqs; int64_t *d_c_off; uint8_t *d_c_alt;
	uint64_t node_cap;    // B-tree nodes per slot
	cudaEvent_t ev_prev_kend;   // ev_kend of the batch enqueued last (cfg.batch_order): the next batch's kernels wait for it
	cudaEvent_t ev_base;  // recorded when the ctx was created: origin of the diagnostics timeline (cs_i_slot_times)
	long long base_host_ns;
	Slot *slots;
};

static void slot_free(Slot *s)
{
	if (s->stream) cudaStreamDestroy(s->stream);
	if (s->stream2) cudaStreamDestroy(s->stream2);
	if (s->stream_out) cudaStreamDestroy(s->stream_out);
	for (int i = 0; i < 8; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
	if (s->ev_pack) cudaEventDestroy(s->ev_pack);
	if (s->ev_kend) cudaEventDestroy(s->ev_kend);
	if (s->ev_fork) cudaEventDestroy(s->ev_fork);
	if (s->ev_join) cudaEventDestroy(s->ev_join);
	for (int i = 0; i < 2; ++i) if (s->ev_r3[i]) cudaEventDestroy(s->ev_r3[i]);
	if (s->ev_done) cudaEventDestroy(s->ev_done);
	if (s->ev_copy) cudaEventDestroy(s->ev_copy);
	if (s->ev_kdone) cudaEventDestroy(s->ev_kdone);
	cudaFreeHost(s->h_packed); cudaFreeHost(s->h_nmask);
	cudaFreeHost(s->h_cmems); cudaFreeHost(s->h_rlo); cudaFreeHost(s->h_rhi);
	cudaFree(s->d_cmems); cudaFree(s->d_rlo); cudaFree(s->d_rhi);
	cudaFree(s->d_s_next); cudaFree(s->d_s_qb_len); cudaFree(s->d_order); cudaFree(s->d_klist); cudaFree(s->d_node_cnt); cudaFree(s->d_node_off); cudaFree(s->d_nodes);
	cudaFree(s->d_n_chain); cudaFree(s->d_n_cseed); cudaFree(s->d_l_rep); cudaFree(s->d_chain_off); cudaFree(s->d_cseed_off); cudaFree(s->d_chain_tmp);
	cudaFree(s->d_chains); cudaFree(s->d_cs_lo); cudaFree(s->d_cs_hi); cudaFree(s->d_cs_qbeg); cudaFree(s->d_cs_len);
	cudaFreeHost(s->h_chain_off); cudaFreeHost(s->h_cseed_off); cudaFreeHost(s->h_chains); cudaFreeHost(s->h_cs_lo); cudaFreeHost(s->h_cs_hi);
	cudaFreeHost(s->h_cs_qbeg); cudaFreeHost(s->h_cs_len);
	cudaFreeHost(s->h_bases); cudaFreeHost(s->h_off); cudaFreeHost(s->h_mem_off); cudaFreeHost(s->h_seed_off);
	cudaFreeHost(s->h_mems); cudaFreeHost(s->h_rbeg); cudaFreeHost(s->h_ctrl);
	cudaFree(s->d_bases); cudaFree(s->d_off); cudaFree(s->d_packed); cudaFree(s->d_nmask); cudaFree(s->d_defer_q); cudaFree(s->d_read_last_q); cudaFree(s->d_x_n); cudaFree(s->d_defer_bits); cudaFree(s->d_defer_lx); cudaFree(s->d_thread_lx); cudaFree(s->d_walk_order); cudaFree(s->d_lit_q); cudaFree(s->d_x_off); cudaFree(s->d_stage); cudaFree(s->d_ctrl); cudaFree(s->d_thread_mems); cudaFree(s->d_spill);
	cudaFree(s->d_pool); cudaFree(s->d_mems); cudaFree(s->d_read_pool_off); cudaFree(s->d_read_n_mems);
	cudaFree(s->d_r3_mems); cudaFree(s->d_r3_n_mems); cudaFree(s->d_tot_n_mems);
	cudaFree(s->d_mem_off); cudaFree(s->d_read_n_seeds); cudaFree(s->d_seed_off); cudaFree(s->d_rows); cudaFree(s->d_scan_tmp);
}

extern "C" void cs_ctx_free(cs_ctx_t *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	for (int i = 0; i < ctx->n_slots; ++i) slot_free(&ctx->slots[i]);
	cudaFree(ctx->d_c_off); cudaFree(ctx->d_c_alt);
	if (ctx->ev_base) cudaEventDestroy(ctx->ev_base);
	free(ctx->slots); free(ctx);
}

extern "C" cs_ctx_t *cs_ctx_create(const cs_index_t *idx, uint32_t max_reads, uint64_t max_bases, uint32_t max_read_len,
                                   uint64_t max_mems, uint64_t max_seeds, int n_slots)
{ return cs_ctx_create_ex(idx, max_reads, max_bases, max_read_len, max_mems, max_seeds, n_slots, nullptr); }

extern "C" uint64_t cs_ctx_launches(const cs_ctx_t *ctx) { return ctx ? ctx->n_launch : 0; }

extern "C" int cs_ctx_need(const cs_ctx_t *ctx, int slot, uint64_t *need_mems, uint64_t *need_seeds)
{
	if (!ctx || slot < 0 || slot >= ctx->n_slots) return set_err(CS_E_ARG, "bad ctx / slot");
	if (need_mems) *need_mems = ctx->need_mems[slot];
	if (need_seeds) *need_seeds = ctx->need_seeds[slot];
	return CS_OK;
}

extern "C" cs_ctx_t *cs_ctx_create_ex(const cs_index_t *idx, uint32_t max_reads, uint64_t max_bases, uint32_t max_read_len,
                                      uint64_t max_mems, uint64_t max_seeds, int n_slots, const cs_ctx_config_t *cfg)
{
	cs_ctx *ctx = nullptr;
	if (!idx) { set_err(CS_E_ARG, "null index"); return nullptr; }
	if (max_reads == 0 || max_bases == 0 || max_read_len == 0 || max_read_len > 65535 || n_slots < 1 || n_slots > 16 ||
	    max_bases >= (1ull << 32)) {
		set_err(CS_E_ARG, "bad ctx geometry (reads %u, bases %llu, read_len %u (<= 65535, comp_seed.h:39), slots %d)",
		        max_reads, (unsigned long long)max_bases, max_read_len, n_slots);
		return nullptr;
	}
	if (use_device(idx->device) != CS_OK) return nullptr;
	ctx = (cs_ctx*)calloc(1, sizeof(cs_ctx));
	if (!ctx) { set_err(CS_E_ARG, "out of host memory"); return nullptr; }
	if (cfg) ctx->cfg = *cfg; else cs_ctx_config_default(&ctx->cfg);
	ctx->idx = idx; ctx->device = idx->device; ctx->max_reads = max_reads; ctx->max_bases = max_bases; ctx->max_read_len = max_read_len;
	ctx->max_mems = max_mems ? max_mems : (uint64_t)max_reads * 16;
	ctx->max_seeds = max_seeds ? max_seeds : (uint64_t)max_reads * 32;
	if (ctx->max_mems >= (1ull << 32) || ctx->max_seeds >= (1ull << 32)) { set_err(CS_E_ARG, "result capacities must be < 2^32 per slot"); free(ctx); return nullptr; }
	ctx->n_slots = n_slots;
	ctx->slots = (Slot*)calloc(n_slots, sizeof(Slot));
	{
		const size_t smem = CS_SEED_SMEM_BYTES;
		int per_sm = 0;
		CK(cudaFuncSetAttribute(k_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		CK(cudaFuncSetAttribute(k_seed_long, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed, CS_SEED_BLOCK, smem));
		if (per_sm < 1) { set_err(CS_E_CUDA, "k_seed does not fit on an SM"); goto fail; }
		ctx->grid = idx->n_sm * per_sm;
		uint32_t need = (max_reads + CS_SEED_BLOCK - 1) / CS_SEED_BLOCK;
		if ((uint32_t)ctx->grid > need) ctx->grid = (int)need;
		// the fast kernel needs the 2-bit text + inverse SA, the occurrence filter, the top-of-search table, and
		// reads that fit its shared-memory words; otherwise k_seed takes every read
		ctx->grid_fast = 0;
		{
			const DevIndex &d = idx->d;
			if (ctx->cfg.use_fast != 0 && d.text && d.isa && d.pt && d.pt_k >= 4 && d.pt_k <= 19 && d.kt && d.kt_depth >= 2 &&
			    d.kt_depth < d.pt_k && max_read_len + 32 <= 32 * CS_READ_SMEM) {
				CK(cudaFuncSetAttribute(k_seed_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_FAST_SMEM_BYTES));
				CK(cudaFuncSetAttribute(k_seed_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_FAST_SMEM_BYTES));
				CK(cudaFuncSetAttribute(k_seed_r3_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CS_FAST_SMEM_BYTES));
				CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed_fast, CS_FAST_BLOCK, CS_FAST_SMEM_BYTES));
				if (per_sm >= 1) ctx->grid_fast = std::min<int>(idx->n_sm * per_sm, (int)((max_reads + CS_FAST_BLOCK - 1) / CS_FAST_BLOCK));
			}
		}
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_seed_r3, 256, 0));
		if (per_sm < 1) per_sm = 1;
		ctx->grid_r3 = idx->n_sm * per_sm;
		if ((uint32_t)ctx->grid_r3 > need) ctx->grid_r3 = (int)need;
	}
	ctx->mem_cap = std::min<uint32_t>(2 * max_read_len + 16, 4096);
	ctx->spill_cap = max_read_len > CS_LIST_SMEM ? max_read_len - CS_LIST_SMEM + 1 : 1;
	ctx->defer_cap = 2 * max_reads + 4096;   // a batch that defers more calls than this is rerun without the fast kernel
	if (ctx->cfg.defer_cap > 0) ctx->defer_cap = (uint32_t)ctx->cfg.defer_cap;   // (tests force the rerun)
	for (int i = 0; i < n_slots; ++i) {
		Slot *s = &ctx->slots[i];
		const size_t nthreads = std::max((size_t)ctx->grid * CS_SEED_BLOCK, (size_t)ctx->grid_fast * CS_FAST_BLOCK);
		CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking));
		if (ctx->cfg.overlap_streams > 0) CK(cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking));   // (streams share the device's hardware queues: none is created idle)
		for (int e = 0; e < 8; ++e) CK(cudaEventCreate(&s->ev[e]));
		CK(cudaEventCreate(&s->ev_pack));
		CK(cudaEventCreate(&s->ev_kend));
		CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
		for (int e = 0; e < 2; ++e) CK(cudaEventCreate(&s->ev_r3[e]));
		if (ctx->cfg.l2_persist_mb > 0 && idx->d.kt) { // persisting L2 window over the top of the K-mer table (the "hot top-of-search" of north_star)
			cudaDeviceProp prop;
			CK(cudaGetDeviceProperties(&prop, idx->device));
			size_t want = (size_t)ctx->cfg.l2_persist_mb << 20;
			if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
			if (want > (size_t)prop.accessPolicyMaxWindowSize) want = (size_t)prop.accessPolicyMaxWindowSize;
			const size_t kt_bytes = (size_t)(kt_offset_host(idx->d.kt_depth + 1)) * sizeof(uint4);
			if (want > kt_bytes) want = kt_bytes;
			if (want > 0) {
				CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
				cudaStreamAttrValue av;
				memset(&av, 0, sizeof av);
				av.accessPolicyWindow.base_ptr = (void*)idx->d.kt;
				av.accessPolicyWindow.num_bytes = want;
				av.accessPolicyWindow.hitRatio = 1.0f;
				av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
				av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
				CK(cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av));
				if (s->stream2) CK(cudaStreamSetAttribute(s->stream2, cudaStreamAttributeAccessPolicyWindow, &av));
			}
		}
		CK(cudaEventCreate(&s->ev_done));
		CK(cudaEventCreate(&s->ev_copy));
		CK(cudaEventCreateWithFlags(&s->ev_kdone, cudaEventDisableTiming));
		CK(cudaMallocHost(&s->h_bases, max_bases));
		CK(cudaMallocHost(&s->h_off, ((size_t)max_reads + 1) * 4));
		CK(cudaMallocHost(&s->h_ctrl, sizeof(Ctrl)));
		CK(cudaMalloc(&s->d_bases, max_bases + 256));   // k_pack_reads reads whole aligned words past a read's last base
		CK(cudaMalloc(&s->d_off, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_packed, ((max_bases >> 5) + 2 * (size_t)max_reads + 8) * 8));
		CK(cudaMalloc(&s->d_nmask, ((max_bases >> 5) + 2 * (size_t)max_reads + 8) * 4));
		CK(cudaMalloc(&s->d_ctrl, sizeof(Ctrl)));
		CK(cudaMalloc(&s->d_defer_q, (size_t)ctx->defer_cap * sizeof(uint4)));
		CK(cudaMalloc(&s->d_x_off, (size_t)ctx->defer_cap * 8));
		CK(cudaMalloc(&s->d_x_n, (size_t)ctx->defer_cap * 4));
		CK(cudaMalloc(&s->d_defer_bits, (size_t)ctx->defer_cap * 4));
		CK(cudaMalloc(&s->d_defer_lx, (size_t)ctx->defer_cap * sizeof(uint4)));
		CK(cudaMalloc(&s->d_walk_order, (size_t)ctx->defer_cap * 4));
		CK(cudaMalloc(&s->d_lit_q, (size_t)ctx->defer_cap * 4));
		CK(cudaMalloc(&s->d_read_last_q, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_stage, ctx->max_mems * sizeof(cs_mem_t)));
		CK(cudaMalloc(&s->d_thread_mems, nthreads * ctx->mem_cap * sizeof(cs_mem_t)));
		CK(cudaMalloc(&s->d_thread_lx, nthreads * sizeof(uint4)));
		CK(cudaMalloc(&s->d_spill, nthreads * ctx->spill_cap * sizeof(uint4)));
		CK(cudaMalloc(&s->d_pool, ctx->max_mems * sizeof(cs_mem_t)));
		CK(cudaMalloc(&s->d_mems, ctx->max_mems * sizeof(cs_mem_t)));
		CK(cudaMalloc(&s->d_read_pool_off, (size_t)max_reads * 8));
		CK(cudaMalloc(&s->d_read_n_mems, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_mem_off, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_read_n_seeds, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_seed_off, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_rows, ctx->max_seeds * 8));
		if (ctx->cfg.compact_results > 0) {
			CK(cudaMalloc(&s->d_cmems, ctx->max_mems * sizeof(cs_cmem_t)));
			CK(cudaMalloc(&s->d_rlo, ctx->max_seeds * 4));
			CK(cudaMalloc(&s->d_rhi, ctx->max_seeds));
		}
		s->r3_cap = max_bases / 16 + max_reads + 1;   // enough for min_seed_len >= 15; grown on demand in enqueue_run
		CK(cudaMalloc(&s->d_r3_mems, s->r3_cap * sizeof(cs_mem_t)));
		CK(cudaMalloc(&s->d_r3_n_mems, ((size_t)max_reads + 1) * 4));
		CK(cudaMalloc(&s->d_tot_n_mems, ((size_t)max_reads + 1) * 4));
		s->scan_tmp_bytes = 0;
		CK(cub::DeviceScan::ExclusiveSum(nullptr, s->scan_tmp_bytes, s->d_read_n_mems, s->d_mem_off, (int)max_reads + 1, s->stream));
		CK(cudaMalloc(&s->d_scan_tmp, s->scan_tmp_bytes + 256));
	}
	CK(cudaEventCreate(&ctx->ev_base));
	CK(cudaEventRecord(ctx->ev_base, ctx->slots[0].stream));
	CK(cudaEventSynchronize(ctx->ev_base));
	ctx->base_host_ns = (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
	return ctx;
fail:
	cs_ctx_free(ctx);
	return nullptr;
}

static int check_slot(cs_ctx *ctx, int slot)
{
	if (!ctx) return set_err(CS_E_ARG, "null ctx");
	if (slot < 0 || slot >= ctx->n_slots) return set_err(CS_E_ARG, "slot %d out of range (%d slots)", slot, ctx->n_slots);
	return CS_OK;
}

// enqueue everything that runs on the device for one batch whose inputs are already in d_bases/d_off
static int enqueue_run(cs_ctx *ctx, Slot *s, const cs_seed_opt_t *opt, bool allow_fast = true)
{
	const cs_index *idx = ctx->idx;
	const uint32_t n = s->n_reads;
	const size_t smem = CS_SEED_SMEM_BYTES;
	int grid = std::min<int>(ctx->grid, (int)((n + CS_SEED_BLOCK - 1) / CS_SEED_BLOCK));
	if (grid < 1) grid = 1;
	SeedArgs a;
	CollectArgs c;
	const bool pass3 = opt->max_mem_intv > 0;
	const bool overlap = ctx->cfg.overlap_streams > 0;
	// The third pass depends on k_seed_fast alone (text-assisted: it reads the first-pass SMEMs k_seed_fast left in the
	// pool) or on nothing (k_seed_r3): it runs on the slot's second stream, next to k_seed_walk / k_seed, whose few long
	// tasks leave most of the GPU idle.
	auto launch_r3 = [&](cudaStream_t st, bool text_assisted) -> cudaError_t {
		if (text_assisted) {
			int gf = std::min<int>(ctx->grid_fast, (int)((n + CS_FAST_BLOCK - 1) / CS_FAST_BLOCK));
			k_seed_r3_fast<<<gf < 1 ? 1 : gf, CS_FAST_BLOCK, CS_FAST_SMEM_BYTES, st>>>(idx->d, a);
		} else {
			int g3 = std::min<int>(ctx->grid_r3, (int)((n + 255) / 256));
			if (s->packed_input) { // the general third-pass kernel reads the byte form
				k_unpack_reads<<<(int)std::min<uint64_t>(((uint64_t)n * 8 + 255) / 256, (uint64_t)idx->n_sm * 16), 256, 0, st>>>(
					s->d_packed, s->d_nmask, s->d_off, n, s->d_bases, s->off_bias);
				++ctx->n_launch;
			}
			k_seed_r3<<<g3 < 1 ? 1 : g3, 256, 0, st>>>(idx->d, a);
		}
		++ctx->n_launch;
		return cudaGetLastError();
	};
	bool r3_done = !pass3;
	auto fork_r3 = [&](bool text_assisted) -> cudaError_t { // after the kernel the third pass depends on has been enqueued on s->stream
		cudaError_t e;
		if ((e = cudaEventRecord(s->ev_fork, s->stream)) != cudaSuccess) return e;
		if ((e = cudaStreamWaitEvent(s->stream2, s->ev_fork, 0)) != cudaSuccess) return e;
		if ((e = cudaEventRecord(s->ev_r3[0], s->stream2)) != cudaSuccess) return e;
		if ((e = launch_r3(s->stream2, text_assisted)) != cudaSuccess) return e;
		if ((e = cudaEventRecord(s->ev_r3[1], s->stream2)) != cudaSuccess) return e;
		if ((e = cudaEventRecord(s->ev_join, s->stream2)) != cudaSuccess) return e;
		r3_done = true;
		return cudaSuccess;
	};
	s->opt = *opt;
	if (pass3) { // third-pass seeds of read r live at off[r]/(k+1) + r: at most bases/(k+1) + n entries
		uint64_t need3 = (uint64_t)s->h_off[n] / ((uint32_t)opt->min_seed_len + 1) + n + 1;
		if (need3 > s->r3_cap) {
			CK(cudaStreamSynchronize(s->stream));
			CK(cudaFree(s->d_r3_mems)); s->d_r3_mems = nullptr; s->r3_cap = 0;
			CK(cudaMalloc(&s->d_r3_mems, need3 * sizeof(cs_mem_t)));
			s->r3_cap = need3;
		}
	}
	CK(cudaMemsetAsync(s->d_ctrl, 0, sizeof(Ctrl), s->stream));
	if (ctx->cfg.batch_order != 0 && ctx->ev_prev_kend && ctx->ev_prev_kend != s->ev_kend)   // (input copies above are not held back)
		CK(cudaStreamWaitEvent(s->stream, ctx->ev_prev_kend, 0));
	CK(cudaEventRecord(s->ev[1], s->stream));
	if (!s->packed_input) {
		k_pack_reads<<<(int)std::min<uint64_t>(((uint64_t)n * 8 + 255) / 256, (uint64_t)idx->n_sm * 16), 256, 0, s->stream>>>(
			s->d_bases, s->d_off, n, s->d_packed, s->d_nmask);
		CK(cudaGetLastError()); ++ctx->n_launch;
	}
	CK(cudaEventRecord(s->ev_pack, s->stream));
	a.bases = s->d_bases; a.off = s->d_off; a.n_reads = n; a.opt = *opt;
	a.packed = s->d_packed; a.nmask = s->d_nmask; a.off_bias = s->packed_input ? s->off_bias : 0u;
	a.next_read = s->d_ctrl->next_read;
	a.defer_q = nullptr; a.defer_cap = ctx->defer_cap; a.n_defer = &s->d_ctrl->n_defer;
	a.read_last_q = s->d_read_last_q; a.x_off = s->d_x_off; a.x_n = s->d_x_n; a.defer_bits = s->d_defer_bits; a.defer_lx = s->d_defer_lx; a.thread_lx = s->d_thread_lx; a.lit_q = s->d_lit_q; a.n_lit = &s->d_ctrl->n_lit; a.n_defer_fast = &s->d_ctrl->n_defer_fast; a.walk_order = s->d_walk_order; a.n_walk = &s->d_ctrl->n_walk;
	s->used_fast = false;
	a.thread_mems = s->d_thread_mems; a.mem_cap = ctx->mem_cap;
	a.spill = s->d_spill; a.spill_cap = ctx->spill_cap;
	a.pool = s->d_pool; a.pool_cap = ctx->max_mems; a.pool_used = &s->d_ctrl->pool_used;
	a.read_pool_off = s->d_read_pool_off; a.read_n_mems = s->d_read_n_mems;
	a.r3_mems = s->d_r3_mems; a.r3_n_mems = s->d_r3_n_mems;
	a.counters = s->d_ctrl->counters; a.req = s->d_ctrl->req; a.error = &s->d_ctrl->error;
	// passes 1-2: the call-synchronous fast kernel first; the reads it cannot prove simple (or all reads, when it
	// is not available, or when min_seed_len is below the filter's K) go through the literal kernel
	if (allow_fast && ctx->grid_fast > 0 && opt->min_seed_len >= (int)idx->d.pt_k) {
		int gf = std::min<int>(ctx->grid_fast, (int)((n + CS_FAST_BLOCK - 1) / CS_FAST_BLOCK));
		a.defer_q = s->d_defer_q;
		s->used_fast = true;
		k_seed_fast<<<gf < 1 ? 1 : gf, CS_FAST_BLOCK, CS_FAST_SMEM_BYTES, s->stream>>>(idx->d, a);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaEventRecord(s->ev[6], s->stream));
		CK(cudaMemcpyAsync(&s->d_ctrl->n_defer_fast, &s->d_ctrl->n_defer, 4, cudaMemcpyDeviceToDevice, s->stream));
		if (pass3 && overlap) CK(fork_r3(ctx->cfg.use_r3_fast != 0));
		{ // the walk tasks, longest first (counting sort: count, scan, scatter)
			const int go = std::max(1, std::min<int>(idx->n_sm * 8, (int)(((uint64_t)n + 2047) / 2048)));
			k_walk_count<<<go, 256, 0, s->stream>>>(idx->d, a, s->d_ctrl->walk_hist);
			k_walk_scan<<<1, 32, 0, s->stream>>>(s->d_ctrl->walk_hist, s->d_ctrl->walk_cursor, &s->d_ctrl->n_walk);
			k_walk_scatter<<<go, 256, 0, s->stream>>>(idx->d, a, s->d_ctrl->walk_cursor, s->d_walk_order);
			CK(cudaGetLastError()); ctx->n_launch += 3;
		}
		k_seed_walk<<<gf < 1 ? 1 : gf, CS_FAST_BLOCK, CS_FAST_SMEM_BYTES, s->stream>>>(idx->d, a);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaEventRecord(s->ev[7], s->stream));
	} else {
		CK(cudaEventRecord(s->ev[6], s->stream));
		CK(cudaEventRecord(s->ev[7], s->stream));
		if (pass3 && overlap) CK(fork_r3(false));
	}
	// (the spill stride inside k_seed follows the launched grid)
	if (s->used_fast && ctx->cfg.lit_ctas_per_sm > 0) // call mode: few, long tasks -- fewer resident warps make each trip of the state machine faster
		grid = std::min<int>(grid, idx->n_sm * ctx->cfg.lit_ctas_per_sm);
	if (ctx->max_read_len + 32 <= 32 * CS_READ_SMEM) k_seed<<<grid, CS_SEED_BLOCK, smem, s->stream>>>(idx->d, a);
	else k_seed_long<<<grid, CS_SEED_BLOCK, smem, s->stream>>>(idx->d, a);
	CK(cudaGetLastError()); ++ctx->n_launch;
	CK(cudaEventRecord(s->ev[5], s->stream));
	if (!r3_done) { // same stream, after the passes-1-2 kernels
		CK(cudaEventRecord(s->ev_r3[0], s->stream));
		CK(launch_r3(s->stream, s->used_fast && ctx->cfg.use_r3_fast != 0));
		CK(cudaEventRecord(s->ev_r3[1], s->stream));
	} else if (pass3) CK(cudaStreamWaitEvent(s->stream, s->ev_join, 0));
	else { CK(cudaEventRecord(s->ev_r3[0], s->stream)); CK(cudaEventRecord(s->ev_r3[1], s->stream)); }
	CK(cudaEventRecord(s->ev[2], s->stream));
	// collect: offsets, sort, SA rows
	k_mem_counts<<<std::min<int>(idx->n_sm * 8, (int)((n + 255) / 256)), 256, 0, s->stream>>>(s->d_read_n_mems, pass3 ? s->d_r3_n_mems : nullptr,
		s->used_fast ? s->d_read_last_q : nullptr, s->d_defer_q, s->d_x_n, n, s->d_tot_n_mems, &s->d_ctrl->tot12, &s->d_ctrl->tot3);
	CK(cudaGetLastError()); ++ctx->n_launch;
	CK(cudaMemsetAsync(s->d_tot_n_mems + n, 0, 4, s->stream));
	CK(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, s->scan_tmp_bytes, s->d_tot_n_mems, s->d_mem_off, (int)n + 1, s->stream));
	ctx->n_launch += 2;   // cub: init + scan kernel
	c.n_reads = n; c.opt = *opt; c.pool = s->d_pool; c.read_pool_off = s->d_read_pool_off; c.read_n_mems = s->d_read_n_mems;
	c.off = s->d_off; c.r3_mems = s->d_r3_mems; c.r3_n_mems = pass3 ? s->d_r3_n_mems : nullptr;
	c.read_last_q = s->used_fast ? s->d_read_last_q : nullptr; c.defer_q = s->d_defer_q; c.x_off = s->d_x_off; c.x_n = s->d_x_n; c.stage = s->d_stage;
	c.mem_off = s->d_mem_off; c.mems = s->d_mems; c.mems_cap = ctx->max_mems; c.read_n_seeds = s->d_read_n_seeds; c.seed_off = s->d_seed_off;
	c.seed_rows = s->d_rows; c.seed_cap = ctx->max_seeds; c.tot_seeds = &s->d_ctrl->tot_seeds; c.error = &s->d_ctrl->error;
	{
		int cgrid = (int)std::min<uint64_t>(((uint64_t)n * 8 + 255) / 256, (uint64_t)idx->n_sm * 16);
		k_collect_sort<<<cgrid, 256, 0, s->stream>>>(c);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaMemsetAsync(s->d_read_n_seeds + n, 0, 4, s->stream));
		CK(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, s->scan_tmp_bytes, s->d_read_n_seeds, s->d_seed_off, (int)n + 1, s->stream));
		ctx->n_launch += 2;
		k_collect_rows<<<cgrid, 256, 0, s->stream>>>(c);
		CK(cudaGetLastError()); ++ctx->n_launch;
	}
	CK(cudaMemcpyAsync(&s->d_ctrl->n_mems, s->d_mem_off + n, 4, cudaMemcpyDeviceToDevice, s->stream));
	CK(cudaMemcpyAsync(&s->d_ctrl->n_seeds, s->d_seed_off + n, 4, cudaMemcpyDeviceToDevice, s->stream));
	CK(cudaEventRecord(s->ev[3], s->stream));
	// SA resolution: the kernel reads n_seeds from ctrl on the device, so nothing waits for the host
	k_sa_resolve<<<idx->n_sm * 8, 256, 0, s->stream>>>(idx->d, &s->d_ctrl->n_seeds, ctx->max_seeds, s->d_rows,
	                                                    &s->d_ctrl->sa_work, &s->d_ctrl->lf_steps);
	CK(cudaGetLastError()); ++ctx->n_launch;
	CK(cudaEventRecord(s->ev[4], s->stream));
	s->chained = false;
	if (ctx->chaining) { // mem_chain + mem_chain_flt on the device (cs_chain.cu): node regions, build + filter, offsets, emit
		ChainArgs ca;
		const int g = std::min<int>(idx->n_sm * 8, (int)((n + 255) / 256));
		k_chain_node_counts<<<g, 256, 0, s->stream>>>(s->d_read_n_seeds, n, s->d_node_cnt);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaMemsetAsync(s->d_node_cnt + n, 0, 4, s->stream));
		CK(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, s->scan_tmp_bytes, s->d_node_cnt, s->d_node_off, (int)n + 1, s->stream));
		ctx->n_launch += 2;
		CK(cudaMemcpyAsync(&s->d_ctrl->n_nodes, s->d_node_off + n, 4, cudaMemcpyDeviceToDevice, s->stream));
		ca.n_reads = n; ca.opt = *opt; ca.copt = ctx->copt; ca.off = s->d_off; ca.mem_off = s->d_mem_off; ca.mems = s->d_mems;
		ca.seed_off = s->d_seed_off; ca.rbeg = s->d_rows; ca.l_pac = ctx->l_pac; ca.n_seqs = ctx->n_seqs; ca.c_off = ctx->d_c_off; ca.c_alt = ctx->d_c_alt;
		ca.s_next = s->d_s_next; ca.s_qb_len = s->d_s_qb_len; ca.chains = s->d_chain_tmp; ca.order = s->d_order; ca.klist = s->d_klist;
		ca.node_off = s->d_node_off; ca.nodes = s->d_nodes; ca.node_cap = ctx->node_cap; ca.seed_cap = ctx->max_seeds; ca.mems_cap = ctx->max_mems;
		ca.n_chain = s->d_n_chain; ca.n_cseed = s->d_n_cseed; ca.l_rep = s->d_l_rep;
		ca.work = &s->d_ctrl->chain_work; ca.error = &s->d_ctrl->error;
		k_chain_build<<<idx->n_sm * 16, 128, 0, s->stream>>>(ca);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaMemsetAsync(s->d_n_chain + n, 0, 4, s->stream));
		CK(cudaMemsetAsync(s->d_n_cseed + n, 0, 4, s->stream));
		CK(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, s->scan_tmp_bytes, s->d_n_chain, s->d_chain_off, (int)n + 1, s->stream));
		CK(cub::DeviceScan::ExclusiveSum(s->d_scan_tmp, s->scan_tmp_bytes, s->d_n_cseed, s->d_cseed_off, (int)n + 1, s->stream));
		ctx->n_launch += 4;
		k_chain_emit<<<(int)std::min<uint64_t>(((uint64_t)n * 8 + 255) / 256, (uint64_t)idx->n_sm * 16), 256, 0, s->stream>>>(
			ca, s->d_chain_off, s->d_cseed_off, ctx->max_mems, ctx->max_seeds, s->d_chains, s->d_cs_lo, s->d_cs_hi, s->d_cs_qbeg, s->d_cs_len);
		CK(cudaGetLastError()); ++ctx->n_launch;
		CK(cudaMemcpyAsync(&s->d_ctrl->n_chains, s->d_chain_off + n, 4, cudaMemcpyDeviceToDevice, s->stream));
		CK(cudaMemcpyAsync(&s->d_ctrl->n_cseeds, s->d_cseed_off + n, 4, cudaMemcpyDeviceToDevice, s->stream));
		s->chained = true;
	}
	if (s->d_cmems) {
		k_compact_results<<<idx->n_sm * 8, 256, 0, s->stream>>>(&s->d_ctrl->n_mems, ctx->max_mems, s->d_mems, s->d_cmems,
		                                                         &s->d_ctrl->n_seeds, ctx->max_seeds, s->d_rows, s->d_rlo, s->d_rhi);
		CK(cudaGetLastError()); ++ctx->n_launch;
	}
	CK(cudaEventRecord(s->ev_kend, s->stream));
	ctx->ev_prev_kend = s->ev_kend;
	CK(cudaStreamWaitEvent(s->stream_out, s->ev_kend, 0));   // everything that goes to the host goes through the slot's output stream
	CK(cudaMemcpyAsync(s->h_ctrl, s->d_ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, s->stream_out));
	CK(cudaEventRecord(s->ev_kdone, s->stream_out));
	s->state = 2;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

// status of a finished run: CS_OK, or which kind of overflow (and, for the global kind, what the batch would need)
static int run_status(cs_ctx *ctx, Slot *s)
{
	const Ctrl *h = s->h_ctrl;
	const int slot = (int)(s - ctx->slots);
	const uint64_t tot_mems = h->tot12 + h->tot3;
	if (h->error == CS_E_READ_OVERFLOW)
		return set_err(CS_E_READ_OVERFLOW, "a read of this batch needs more than %u mems or %u interval-list entries: the per-read scratch follows "
		               "max_read_len (%u); larger max_mems / max_seeds do not help", ctx->mem_cap, ctx->spill_cap + CS_LIST_SMEM, ctx->max_read_len);
	if (s->chained && h->pool_used <= ctx->max_mems && tot_mems <= ctx->max_mems && h->tot_seeds <= ctx->max_seeds
	    && (h->n_chains > ctx->max_mems || h->n_cseeds > ctx->max_seeds)) {
		// the mems and seed positions fitted, the chains made of them did not (chain records share max_mems, chain seeds max_seeds)
		ctx->need_mems[slot] = std::max<uint64_t>((uint64_t)h->n_chains + h->n_chains / 16 + 64, ctx->max_mems);
		ctx->need_seeds[slot] = std::max<uint64_t>((uint64_t)h->n_cseeds + h->n_cseeds / 16 + 64, ctx->max_seeds);
		return set_err(CS_E_OVERFLOW, "chain buffers too small for this batch: %u chains of %llu (max_mems), %u chain seeds of %llu (max_seeds); "
		               "re-create the ctx with the capacities cs_ctx_need reports (%llu, %llu)", h->n_chains, (unsigned long long)ctx->max_mems,
		               h->n_cseeds, (unsigned long long)ctx->max_seeds, (unsigned long long)ctx->need_mems[slot], (unsigned long long)ctx->need_seeds[slot]);
	}
	if (h->error != 0 || h->pool_used > ctx->max_mems || tot_mems > ctx->max_mems || h->tot_seeds > ctx->max_seeds) {
		// passes 1-2 went through the pool (pool_used counts what was asked for, also past the capacity; tot12 misses the
		// reads that did not fit); the seeds of reads whose mems did not fit in the pool are unknown: scale what is known
		uint64_t nm = std::max<uint64_t>(h->pool_used, h->tot12) + h->tot3, ns = h->tot_seeds;
		if (h->pool_used > ctx->max_mems) ns = std::max<uint64_t>(ns, (uint64_t)((double)ctx->max_seeds * ((double)nm / (double)ctx->max_mems) * 1.25));
		ctx->need_mems[slot] = std::max<uint64_t>(nm + nm / 16 + 64, ctx->max_mems);
		ctx->need_seeds[slot] = std::max<uint64_t>(ns + ns / 16 + 64, ctx->max_seeds);
		return set_err(CS_E_OVERFLOW, "result buffers too small for this batch: mems %llu of %llu, seeds %llu of %llu; "
		               "re-create the ctx with the capacities cs_ctx_need reports (%llu, %llu)",
		               (unsigned long long)nm, (unsigned long long)ctx->max_mems, (unsigned long long)ns, (unsigned long long)ctx->max_seeds,
		               (unsigned long long)ctx->need_mems[slot], (unsigned long long)ctx->need_seeds[slot]);
	}
	return CS_OK;
}

// wait for the device side of a run and check its status
static int finish_run(cs_ctx *ctx, Slot *s)
{
	CK(cudaStreamSynchronize(s->stream_out));   // (waits for the kernels on s->stream through ev_kend)
	if (s->used_fast && s->h_ctrl->n_defer > ctx->defer_cap) { // more hard calls than the queue holds (repeat-rich batch): literal kernel alone
		cs_seed_opt_t o = s->opt;
		if (enqueue_run(ctx, s, &o, false) != CS_OK) return CS_E_CUDA;
		CK(cudaStreamSynchronize(s->stream_out));
	}
	s->state = 3;
	return run_status(ctx, s);
fail:
	return CS_E_CUDA;
}

static void fill_result(cs_ctx *ctx, Slot *s, cs_result_t *out, bool host_ptrs)
{
	memset(out, 0, sizeof *out);
	out->n_reads = s->n_reads;
	out->n_mems = s->h_ctrl->n_mems; out->n_seeds = s->h_ctrl->n_seeds;
	out->counters.ext_queries = s->h_ctrl->counters[0];
	out->counters.ext_calls = s->h_ctrl->counters[1];
	out->counters.sal_queries = s->h_ctrl->n_seeds;
	out->counters.sal_calls = s->h_ctrl->lf_steps;
	if (host_ptrs) { out->mem_off = s->h_mem_off; out->mems = s->h_mems; out->seed_off = s->h_seed_off; out->rbeg = s->h_rbeg; }
	cudaEventElapsedTime(&out->kernel_ms[0], s->ev[1], s->ev[2]);
	cudaEventElapsedTime(&out->kernel_ms[1], s->ev[2], s->ev[3]);
	cudaEventElapsedTime(&out->kernel_ms[2], s->ev[3], s->ev[4]);
	cudaEventElapsedTime(&out->kernel_ms[3], s->ev[0], s->ev_done);
	cudaEventElapsedTime(&out->kernel_ms[4], s->ev[1], s->ev[5]);
	cudaEventElapsedTime(&out->kernel_ms[5], s->ev[5], s->ev[2]);
	cudaEventElapsedTime(&out->kernel_ms[6], s->ev[1], s->ev[6]);
	cudaEventElapsedTime(&out->kernel_ms[7], s->ev[6], s->ev[7]);
	cudaEventElapsedTime(&out->kernel_ms2[0], s->ev[7], s->ev[5]);
	cudaEventElapsedTime(&out->kernel_ms2[1], s->ev_r3[0], s->ev_r3[1]);
	cudaEventElapsedTime(&out->kernel_ms2[2], s->ev[1], s->ev_pack);
	out->n_deferred = s->h_ctrl->n_defer;
	for (int k = 0; k < 4; ++k) out->gather_requests[k] = s->h_ctrl->req[k];
	// k_sa_resolve: one 8-byte gather per seed with the dense SA, else one Occ sector per LF step plus the sample
	out->gather_requests[4] = s->h_ctrl->n_seeds + s->h_ctrl->lf_steps;
	(void)ctx;
}

static int check_batch(cs_ctx *ctx, uint32_t n_reads, const uint32_t *offsets)
{
	if (n_reads == 0 || n_reads > ctx->max_reads) return set_err(CS_E_ARG, "n_reads %u outside [1, %u]", n_reads, ctx->max_reads);
	if (offsets[0] != 0) return set_err(CS_E_ARG, "offsets[0] must be 0");
	if (offsets[n_reads] > ctx->max_bases) return set_err(CS_E_ARG, "batch has %u bases, ctx sized for %llu", offsets[n_reads], (unsigned long long)ctx->max_bases);
	for (uint32_t r = 0; r < n_reads; ++r) {
		if (offsets[r + 1] < offsets[r]) return set_err(CS_E_ARG, "offsets not monotone at read %u", r);
		if (offsets[r + 1] - offsets[r] > ctx->max_read_len)
			return set_err(CS_E_ARG, "read %u has length %u > max_read_len %u", r, offsets[r + 1] - offsets[r], ctx->max_read_len);
	}
	return CS_OK;
}

static void prefetch_ready(cs_ctx *ctx, const Slot *except);
static int check_opt(const cs_seed_opt_t *opt);

extern "C" int cs_seed_batch_stage(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!bases || !offsets) return set_err(CS_E_ARG, "null argument");
	if ((rc = check_batch(ctx, n_reads, offsets)) != CS_OK) return rc;
	Slot *s = &ctx->slots[slot];
	if (s->state == 2 || s->state == 4) return set_err(CS_E_STATE, "slot %d is busy", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	if (offsets != s->h_off) memcpy(s->h_off, offsets, ((size_t)n_reads + 1) * 4);
	s->n_reads = n_reads;
	CK(cudaEventRecord(s->ev[0], s->stream));
	{
		// bases already in page-locked memory (cs_host_register / cudaHostAlloc by the caller): DMA straight
		// from the caller's buffer, which must then stay untouched until the slot is waited on;
		// otherwise stage through the slot's own pinned buffer
		cudaPointerAttributes pa;
		bool pinned = cudaPointerGetAttributes(&pa, bases) == cudaSuccess && pa.type == cudaMemoryTypeHost;
		cudaGetLastError();
		if (pinned) CK(cudaMemcpyAsync(s->d_bases, bases, offsets[n_reads], cudaMemcpyHostToDevice, s->stream));
		else {
			memcpy(s->h_bases, bases, offsets[n_reads]);
			CK(cudaMemcpyAsync(s->d_bases, s->h_bases, offsets[n_reads], cudaMemcpyHostToDevice, s->stream));
		}
	}
	CK(cudaMemcpyAsync(s->d_off, s->h_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, s->stream));
	s->packed_input = false;
	s->state = 1;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" uint64_t cs_packed_words(uint32_t n_reads, const uint32_t *offsets)
{
	return offsets ? (uint64_t)(offsets[n_reads] >> 5) + 2ull * n_reads : 0;
}

extern "C" int cs_pack_reads_host(uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets, uint64_t *packed, uint32_t *nmask, int n_threads)
{ // the layout of cs_seed_batch_submit_packed; every word a read owns is written, the words between reads are left "all N"
	if (!bases || !offsets || !packed || !nmask) return set_err(CS_E_ARG, "null argument");
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	const uint64_t nw_tot = cs_packed_words(n_reads, offsets);
	auto work = [=](uint32_t r0, uint32_t r1) {
		for (uint32_t r = r0; r < r1; ++r) {
			const uint32_t o = offsets[r], len = offsets[r + 1] - o;
			const uint64_t w0 = (uint64_t)(o >> 5) + 2ull * r, nw = (uint64_t)(len >> 5) + 2;
			const uint8_t *q = bases + o;
			for (uint64_t w = 0; w < nw; ++w) {
				uint64_t v = 0; uint32_t m = 0xffffffffu;
				const uint32_t p0 = (uint32_t)w << 5, cnt = p0 < len ? (len - p0 < 32 ? len - p0 : 32) : 0;
				for (uint32_t j = 0; j < cnt; ++j) {
					const uint32_t c = q[p0 + j];
					if (c <= 3) { v |= (uint64_t)c << (2 * j); m &= ~(1u << j); }
				}
				packed[w0 + w] = v; nmask[w0 + w] = m;
			}
			// words up to the next read's first word (there are none for equal-length reads that are multiples of 32 apart)
			const uint64_t next0 = r + 1 < n_reads ? (uint64_t)(offsets[r + 1] >> 5) + 2ull * (r + 1) : nw_tot;
			for (uint64_t w = w0 + nw; w < next0; ++w) { packed[w] = 0; nmask[w] = 0xffffffffu; }
		}
	};
	if (n_threads == 1 || n_reads < 4096) { work(0, n_reads); return CS_OK; }
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; ++t) {
		const uint32_t r0 = (uint32_t)((uint64_t)n_reads * t / n_threads), r1 = (uint32_t)((uint64_t)n_reads * (t + 1) / n_threads);
		th.emplace_back(work, r0, r1);
	}
	for (auto &t : th) t.join();
	return CS_OK;
}

// a batch whose packed words are the slice [packed, packed + nw) of a larger packed set: read r of the batch starts at word
// ((off_bias + h_off[r]) >> 5) + 2r of the slice (off_bias = 0 for a batch packed on its own).  s->h_off / s->n_reads are set.
static int submit_packed_words(cs_ctx *ctx, Slot *s, const uint64_t *packed, const uint32_t *nmask, uint64_t nw, uint32_t off_bias, const cs_seed_opt_t *opt)
{
	const size_t cap = ((size_t)(ctx->max_bases >> 5) + 2 * (size_t)ctx->max_reads + 8);
	const uint32_t n_reads = s->n_reads;
	if (nw > cap) return set_err(CS_E_ARG, "packed batch has %llu words, ctx sized for %zu", (unsigned long long)nw, cap);
	CK(cudaEventRecord(s->ev[0], s->stream));
	{
		cudaPointerAttributes pa;
		bool pinned = cudaPointerGetAttributes(&pa, packed) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
		              cudaPointerGetAttributes(&pa, nmask) == cudaSuccess && pa.type == cudaMemoryTypeHost;
		cudaGetLastError();
		const uint64_t *src_p = packed; const uint32_t *src_m = nmask;
		if (!pinned) { // stage through the slot's own page-locked buffers
			if (!s->h_packed) { CK(cudaMallocHost(&s->h_packed, cap * 8)); CK(cudaMallocHost(&s->h_nmask, cap * 4)); }
			memcpy(s->h_packed, packed, nw * 8); memcpy(s->h_nmask, nmask, nw * 4);
			src_p = s->h_packed; src_m = s->h_nmask;
		}
		CK(cudaMemcpyAsync(s->d_packed, src_p, nw * 8, cudaMemcpyHostToDevice, s->stream));
		CK(cudaMemcpyAsync(s->d_nmask, src_m, nw * 4, cudaMemcpyHostToDevice, s->stream));
		CK(cudaMemcpyAsync(s->d_off, s->h_off, ((size_t)n_reads + 1) * 4, cudaMemcpyHostToDevice, s->stream));
	}
	s->packed_input = true; s->off_bias = off_bias;
	s->state = 1;
	s->want_fetch = true;
	{
		const int rc = enqueue_run(ctx, s, opt);
		prefetch_ready(ctx, nullptr);
		return rc;
	}
fail:
	return CS_E_CUDA;
}

extern "C" int cs_seed_batch_submit_packed(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint64_t *packed, const uint32_t *nmask,
                                           const uint32_t *offsets, const cs_seed_opt_t *opt)
{
	int rc;
	if ((rc = check_opt(opt)) != CS_OK) return rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!packed || !nmask || !offsets) return set_err(CS_E_ARG, "null argument");
	if ((rc = check_batch(ctx, n_reads, offsets)) != CS_OK) return rc;
	Slot *s = &ctx->slots[slot];
	if (s->state == 2 || s->state == 4) return set_err(CS_E_STATE, "slot %d is busy", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	memcpy(s->h_off, offsets, ((size_t)n_reads + 1) * 4);
	s->n_reads = n_reads;
	return submit_packed_words(ctx, s, packed, nmask, cs_packed_words(n_reads, offsets), 0, opt);
}

// ---------------------------------------------------------------------------------------------
// internals used by the multi-device pipeline (cs_multi.cu): a batch is reads [r0, r0 + n) of a set described by 64-bit
// offsets; results are copied in the compact wire format straight to where the caller wants them (page-locked memory)
// ---------------------------------------------------------------------------------------------
int cs_i_submit(cs_ctx *ctx, int slot, uint32_t n, const uint64_t *off64, const uint8_t *bases, const uint64_t *packed, const uint32_t *nmask,
                uint64_t r0, const cs_seed_opt_t *opt)
{
	int rc;
	if ((rc = check_opt(opt)) != CS_OK) return rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (n == 0 || n > ctx->max_reads) return set_err(CS_E_ARG, "n_reads %u outside [1, %u]", n, ctx->max_reads);
	Slot *s = &ctx->slots[slot];
	if (s->state == 2 || s->state == 4) return set_err(CS_E_STATE, "slot %d is busy", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	const uint64_t o0 = off64[r0];
	if (off64[r0 + n] - o0 > ctx->max_bases) return set_err(CS_E_ARG, "batch has %llu bases, ctx sized for %llu", (unsigned long long)(off64[r0 + n] - o0), (unsigned long long)ctx->max_bases);
	for (uint32_t r = 0; r <= n; ++r) s->h_off[r] = (uint32_t)(off64[r0 + r] - o0);
	for (uint32_t r = 0; r < n; ++r)
		if (s->h_off[r + 1] < s->h_off[r] || s->h_off[r + 1] - s->h_off[r] > ctx->max_read_len)
			return set_err(CS_E_ARG, "read %llu has length %u > max_read_len %u (or offsets not monotone)", (unsigned long long)(r0 + r), s->h_off[r + 1] - s->h_off[r], ctx->max_read_len);
	s->n_reads = n;
	if (packed) { // slice of the set-global packed arrays (cs_pack_reads_host64): words [(o0 >> 5) + 2 r0, (o1 >> 5) + 2 (r0 + n))
		const uint64_t w_lo = (o0 >> 5) + 2 * r0, w_hi = (off64[r0 + n] >> 5) + 2 * (r0 + n);
		return submit_packed_words(ctx, s, packed + w_lo, nmask + w_lo, w_hi - w_lo, (uint32_t)(o0 & 31), opt);
	}
	if ((rc = cs_seed_batch_stage(ctx, slot, n, bases + o0, s->h_off)) != CS_OK) return rc;
	s->want_fetch = true;
	rc = enqueue_run(ctx, s, opt);
	return rc;
}

// kernels of the slot done: status and result sizes
int cs_i_finish(cs_ctx *ctx, int slot, uint64_t *n_mems, uint64_t *n_seeds)
{
	Slot *s = &ctx->slots[slot];
	if (s->state != 2) return set_err(CS_E_STATE, "slot %d has no batch in flight", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	const int rc = finish_run(ctx, s);
	if (rc != CS_OK) { s->state = 1; return rc; }
	if (s->chained) {
		*n_mems = s->h_ctrl->n_chains; *n_seeds = s->h_ctrl->n_cseeds;
		if (*n_mems > ctx->max_mems || *n_seeds > ctx->max_seeds) return set_err(CS_E_OVERFLOW, "chain buffers too small for this batch");
	} else { *n_mems = s->h_ctrl->n_mems; *n_seeds = s->h_ctrl->n_seeds; }
	return CS_OK;
}

int cs_i_fetch_chains_into(cs_ctx *ctx, int slot, uint32_t *chain_off, uint32_t *cseed_off, cs_chain_t *ch, uint32_t *lo, uint8_t *hi, uint16_t *qb, uint16_t *ln)
{
	Slot *s = &ctx->slots[slot];
	const uint32_t n = s->n_reads;
	if (s->state != 3 || !s->chained) return set_err(CS_E_STATE, "slot %d has no finished chained batch", slot);
	CK(cudaEventRecord(s->ev_copy, s->stream_out));
	CK(cudaMemcpyAsync(chain_off, s->d_chain_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	CK(cudaMemcpyAsync(cseed_off, s->d_cseed_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	if (s->h_ctrl->n_chains) CK(cudaMemcpyAsync(ch, s->d_chains, (size_t)s->h_ctrl->n_chains * sizeof(cs_chain_t), cudaMemcpyDeviceToHost, s->stream_out));
	if (s->h_ctrl->n_cseeds) {
		const size_t k = s->h_ctrl->n_cseeds;
		CK(cudaMemcpyAsync(lo, s->d_cs_lo, k * 4, cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(hi, s->d_cs_hi, k, cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(qb, s->d_cs_qbeg, k * 2, cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(ln, s->d_cs_len, k * 2, cudaMemcpyDeviceToHost, s->stream_out));
	}
	CK(cudaEventRecord(s->ev_done, s->stream_out));
	s->state = 4;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

// enqueue the result copies of a finished slot (compact wire format) to caller-owned page-locked memory
int cs_i_fetch_compact_into(cs_ctx *ctx, int slot, uint32_t *mem_off, uint32_t *seed_off, cs_cmem_t *cm, uint32_t *lo, uint8_t *hi)
{
	Slot *s = &ctx->slots[slot];
	const uint32_t n = s->n_reads;
	if (s->state != 3 || !s->d_cmems) return set_err(CS_E_STATE, "slot %d has no finished batch with compact results", slot);
	CK(cudaEventRecord(s->ev_copy, s->stream_out));
	CK(cudaMemcpyAsync(mem_off, s->d_mem_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	CK(cudaMemcpyAsync(seed_off, s->d_seed_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	if (s->h_ctrl->n_mems) CK(cudaMemcpyAsync(cm, s->d_cmems, (size_t)s->h_ctrl->n_mems * sizeof(cs_cmem_t), cudaMemcpyDeviceToHost, s->stream_out));
	if (s->h_ctrl->n_seeds) {
		CK(cudaMemcpyAsync(lo, s->d_rlo, (size_t)s->h_ctrl->n_seeds * 4, cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(hi, s->d_rhi, (size_t)s->h_ctrl->n_seeds, cudaMemcpyDeviceToHost, s->stream_out));
	}
	CK(cudaEventRecord(s->ev_done, s->stream_out));
	s->state = 4;
	return CS_OK;
fail:
	return CS_E_CUDA;
}

int cs_i_fetch_wait(cs_ctx *ctx, int slot, cs_counters_t *cnt, float *slot_ms)
{
	Slot *s = &ctx->slots[slot];
	if (s->state != 4) return set_err(CS_E_STATE, "slot %d has no copy in flight", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	CK(cudaEventSynchronize(s->ev_done));
	s->state = 3;
	if (cnt) {
		cnt->ext_queries += s->h_ctrl->counters[0]; cnt->ext_calls += s->h_ctrl->counters[1];
		cnt->sal_queries += s->h_ctrl->n_seeds; cnt->sal_calls += s->h_ctrl->lf_steps;
	}
	if (slot_ms) { // [0] submit -> kernels start (H2D and whatever ran before on the GPU), [1] seeding kernels, [2] collect + SA (+ chaining, compaction), [3] kernels end -> results on the host
		cudaEventElapsedTime(slot_ms + 0, s->ev[0], s->ev[1]);
		cudaEventElapsedTime(slot_ms + 1, s->ev[1], s->ev[2]);
		cudaEventElapsedTime(slot_ms + 2, s->ev[2], s->ev_kend);
		cudaEventElapsedTime(slot_ms + 3, s->ev_kend, s->ev_done);
	}
	return CS_OK;
fail:
	return CS_E_CUDA;
}

// diagnostics timeline of the batch last fetched from the slot: ms since the ctx was created of [0] submit (input copy enqueued),
// [1] first kernel, [2] last kernel done, [3] result copies enqueued, [4] results on the host; *base_host_ns = host clock at the origin
int cs_i_slot_times(cs_ctx *ctx, int slot, float *t5, long long *base_host_ns)
{
	Slot *s = &ctx->slots[slot];
	cudaEvent_t e[5] = { s->ev[0], s->ev[1], s->ev_kend, s->ev_copy, s->ev_done };
	for (int k = 0; k < 5; ++k) if (cudaEventElapsedTime(t5 + k, ctx->ev_base, e[k]) != cudaSuccess) { cudaGetLastError(); t5[k] = -1.f; }
	if (base_host_ns) *base_host_ns = ctx->base_host_ns;
	return CS_OK;
}

// non-blocking: 1 if the kernels (state 2) / the result copies (state 4) of the slot have finished, 0 if not yet, < 0 on error
int cs_i_poll(cs_ctx *ctx, int slot)
{
	Slot *s = &ctx->slots[slot];
	if (s->state != 2 && s->state != 4) return set_err(CS_E_STATE, "slot %d has nothing in flight", slot);
	const cudaError_t e = cudaEventQuery(s->state == 2 ? s->ev_kdone : s->ev_done);
	if (e == cudaSuccess) return 1;
	if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
	return set_err(CS_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
}

void cs_i_ctx_caps(const cs_ctx *ctx, uint64_t *max_mems, uint64_t *max_seeds) { *max_mems = ctx->max_mems; *max_seeds = ctx->max_seeds; }
const cs_index *cs_i_ctx_index(const cs_ctx *ctx) { return ctx->idx; }

static int check_opt(const cs_seed_opt_t *opt)
{
	if (!opt) return set_err(CS_E_ARG, "null options");
	if (opt->min_seed_len < 1 || opt->max_occ < 1 || opt->max_mem_intv < 0 || opt->split_width < 0)
		return set_err(CS_E_ARG, "bad seeding options (k %d, split_len %d, s %d, y %d, c %d)", opt->min_seed_len, opt->split_len,
		               opt->split_width, opt->max_mem_intv, opt->max_occ);
	return CS_OK;
}

extern "C" int cs_seed_batch_run_staged(cs_ctx_t *ctx, int slot, const cs_seed_opt_t *opt)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if ((rc = check_opt(opt)) != CS_OK) return rc;
	Slot *s = &ctx->slots[slot];
	if (s->state == 0) return set_err(CS_E_STATE, "slot %d has no staged batch", slot);
	if (s->state == 2 || s->state == 4) return set_err(CS_E_STATE, "slot %d is busy", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	if (s->state == 3) cudaEventRecord(s->ev[0], s->stream);
	s->want_fetch = false;
	return enqueue_run(ctx, s, opt);
}

extern "C" int cs_seed_batch_submit(cs_ctx_t *ctx, int slot, uint32_t n_reads, const uint8_t *bases, const uint32_t *offsets,
                                    const cs_seed_opt_t *opt)
{
	int rc;
	if ((rc = check_opt(opt)) != CS_OK) return rc;
	if ((rc = cs_seed_batch_stage(ctx, slot, n_reads, bases, offsets)) != CS_OK) return rc;
	ctx->slots[slot].want_fetch = true;
	rc = enqueue_run(ctx, &ctx->slots[slot], opt);
	prefetch_ready(ctx, nullptr);
	return rc;
}

extern "C" int cs_seed_batch_wait_device(cs_ctx_t *ctx, int slot, cs_result_t *out)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!out) return set_err(CS_E_ARG, "null result");
	Slot *s = &ctx->slots[slot];
	if (s->state != 2) return set_err(CS_E_STATE, "slot %d has no batch in flight", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	rc = finish_run(ctx, s);
	cudaEventRecord(s->ev_done, s->stream_out);
	cudaEventSynchronize(s->ev_done);
	if (rc != CS_OK) { s->state = 1; return rc; }
	fill_result(ctx, s, out, false);
	return CS_OK;
}

static int fetch(cs_ctx *ctx, Slot *s, bool compact = false)
{
	const uint32_t n = s->n_reads;
	if (!s->h_mem_off) { // pinned result buffers are allocated on first use (device-resident runs never need them)
		CK(cudaMallocHost(&s->h_mem_off, ((size_t)ctx->max_reads + 1) * 4));
		CK(cudaMallocHost(&s->h_seed_off, ((size_t)ctx->max_reads + 1) * 4));
	}
	if (!compact && !s->h_mems) {
		CK(cudaMallocHost(&s->h_mems, ctx->max_mems * sizeof(cs_mem_t)));
		CK(cudaMallocHost(&s->h_rbeg, ctx->max_seeds * 8));
	}
	if (compact && !s->h_cmems) {
		CK(cudaMallocHost(&s->h_cmems, ctx->max_mems * sizeof(cs_cmem_t)));
		CK(cudaMallocHost(&s->h_rlo, ctx->max_seeds * 4));
		CK(cudaMallocHost(&s->h_rhi, ctx->max_seeds));
	}
	CK(cudaMemcpyAsync(s->h_mem_off, s->d_mem_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	CK(cudaMemcpyAsync(s->h_seed_off, s->d_seed_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
	if (!compact) {
		if (s->h_ctrl->n_mems) CK(cudaMemcpyAsync(s->h_mems, s->d_mems, (size_t)s->h_ctrl->n_mems * sizeof(cs_mem_t), cudaMemcpyDeviceToHost, s->stream_out));
		if (s->h_ctrl->n_seeds) CK(cudaMemcpyAsync(s->h_rbeg, s->d_rows, (size_t)s->h_ctrl->n_seeds * 8, cudaMemcpyDeviceToHost, s->stream_out));
	} else {
		if (s->h_ctrl->n_mems) CK(cudaMemcpyAsync(s->h_cmems, s->d_cmems, (size_t)s->h_ctrl->n_mems * sizeof(cs_cmem_t), cudaMemcpyDeviceToHost, s->stream_out));
		if (s->h_ctrl->n_seeds) {
			CK(cudaMemcpyAsync(s->h_rlo, s->d_rlo, (size_t)s->h_ctrl->n_seeds * 4, cudaMemcpyDeviceToHost, s->stream_out));
			CK(cudaMemcpyAsync(s->h_rhi, s->d_rhi, (size_t)s->h_ctrl->n_seeds, cudaMemcpyDeviceToHost, s->stream_out));
		}
	}
	s->fetched_compact = compact;
	CK(cudaEventRecord(s->ev_done, s->stream_out));
	return CS_OK;
fail:
	return CS_E_CUDA;
}

static bool results_fit(const cs_ctx *ctx, const Slot *s)
{
	const Ctrl *h = s->h_ctrl;
	return !(h->error != 0 || h->pool_used > ctx->max_mems || h->tot12 + h->tot3 > ctx->max_mems || h->tot_seeds > ctx->max_seeds);
}

// Non-blocking: the result copies of every other slot whose kernels have finished are enqueued now, so that the
// device-to-host engine goes from one batch to the next without waiting for the host to ask (the results are the
// larger transfer: 355 bytes per read on ordinary data).  Anything unusual is left to the blocking path.
static void prefetch_ready(cs_ctx *ctx, const Slot *except)
{
	// Off unless cfg.prefetch_results: on the cfg2 bench it made the host-buffer path slower (79 vs 94 M reads/s, same box):
	// two result copies then share the link, the slot being waited on returns later, and the next batch is submitted later.
	if (!ctx->cfg.prefetch_results) return;
	for (int i = 0; i < ctx->n_slots; ++i) {
		Slot *s = &ctx->slots[i];
		if (s == except || s->state != 2 || !s->want_fetch) continue;
		if (cudaEventQuery(s->ev_kdone) != cudaSuccess) { cudaGetLastError(); continue; }
		if ((s->used_fast && s->h_ctrl->n_defer > ctx->defer_cap) || !results_fit(ctx, s)) continue;
		if (fetch(ctx, s) == CS_OK) s->state = 4;   // 4: results on their way to the host
	}
}

// wait for the result copies of a slot; meanwhile keep the copy engine fed with the other slots' results
static int fetch_wait(cs_ctx *ctx, Slot *s)
{
	for (;;) {
		cudaError_t e = cudaEventQuery(s->ev_done);
		if (e == cudaSuccess) return CS_OK;
		if (e != cudaErrorNotReady) return set_err(CS_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
		cudaGetLastError();
		prefetch_ready(ctx, s);
		{ struct timespec ts = {0, 100000}; nanosleep(&ts, nullptr); }   // 0.1 ms: a busy poll slows the driver down
	}
}

extern "C" int cs_seed_batch_fetch(cs_ctx_t *ctx, int slot, cs_result_t *out)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!out) return set_err(CS_E_ARG, "null result");
	Slot *s = &ctx->slots[slot];
	if (s->state != 3) return set_err(CS_E_STATE, "slot %d has no finished batch", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	if ((rc = fetch(ctx, s)) != CS_OK) return rc;
	if ((rc = fetch_wait(ctx, s)) != CS_OK) return rc;
	fill_result(ctx, s, out, true);
	return CS_OK;
}

extern "C" int cs_seed_batch_wait(cs_ctx_t *ctx, int slot, cs_result_t *out)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!out) return set_err(CS_E_ARG, "null result");
	Slot *s = &ctx->slots[slot];
	if (s->state != 2 && s->state != 4) return set_err(CS_E_STATE, "slot %d has no batch in flight", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	if (s->state == 2) {
		if ((rc = finish_run(ctx, s)) != CS_OK) { s->state = 1; return rc; }
		if ((rc = fetch(ctx, s)) != CS_OK) return rc;
	}
	if ((rc = fetch_wait(ctx, s)) != CS_OK) return rc;
	s->state = 3;
	fill_result(ctx, s, out, true);
	return CS_OK;
}

extern "C" int cs_seed_batch_wait_compact(cs_ctx_t *ctx, int slot, cs_compact_result_t *out)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!out) return set_err(CS_E_ARG, "null result");
	Slot *s = &ctx->slots[slot];
	if (!s->d_cmems) return set_err(CS_E_STATE, "the ctx was created without compact_results");
	if (s->state != 2 && s->state != 3) return set_err(CS_E_STATE, "slot %d has no batch in flight or finished", slot);
	{ const int rc_ = use_device(ctx->idx->device); if (rc_ != CS_OK) return rc_; }
	if (s->state == 2 && (rc = finish_run(ctx, s)) != CS_OK) { s->state = 1; return rc; }
	if ((rc = fetch(ctx, s, true)) != CS_OK) return rc;
	if ((rc = fetch_wait(ctx, s)) != CS_OK) return rc;
	s->state = 3;
	memset(out, 0, sizeof *out);
	out->n_reads = s->n_reads; out->n_mems = s->h_ctrl->n_mems; out->n_seeds = s->h_ctrl->n_seeds;
	out->mem_off = s->h_mem_off; out->cmems = s->h_cmems; out->seed_off = s->h_seed_off; out->rbeg_lo = s->h_rlo; out->rbeg_hi = s->h_rhi;
	out->counters.ext_queries = s->h_ctrl->counters[0]; out->counters.ext_calls = s->h_ctrl->counters[1];
	out->counters.sal_queries = s->h_ctrl->n_seeds; out->counters.sal_calls = s->h_ctrl->lf_steps;
	return CS_OK;
}

extern "C" int cs_ctx_set_chaining(cs_ctx_t *ctx, const cs_bns_view_t *bns, const cs_chain_opt_t *opt)
{
	if (!ctx) return set_err(CS_E_ARG, "null ctx");
	{ const int rc_ = use_device(ctx->device); if (rc_ != CS_OK) return rc_; }
	for (int i = 0; i < ctx->n_slots; ++i) if (ctx->slots[i].state == 2 || ctx->slots[i].state == 4) return set_err(CS_E_STATE, "slot %d is busy", i);
	if (!bns) { ctx->chaining = false; return CS_OK; }
	if (!opt || bns->n_seqs < 1 || !bns->offset || bns->l_pac <= 0 || (uint64_t)bns->l_pac * 2 != ctx->idx->d.seq_len)
		return set_err(CS_E_ARG, "bad contig table (n_seqs %d, l_pac %lld; the index has %llu rows)", bns ? bns->n_seqs : 0, bns ? (long long)bns->l_pac : 0, (unsigned long long)ctx->idx->d.seq_len);
	if (ctx->max_read_len > 65535) return set_err(CS_E_ARG, "chaining packs read positions in 16 bits");
	cudaFree(ctx->d_c_off); cudaFree(ctx->d_c_alt); ctx->d_c_off = nullptr; ctx->d_c_alt = nullptr;
	CK(cudaMalloc(&ctx->d_c_off, (size_t)bns->n_seqs * 8));
	CK(cudaMemcpy(ctx->d_c_off, bns->offset, (size_t)bns->n_seqs * 8, cudaMemcpyHostToDevice));
	if (bns->is_alt) {
		CK(cudaMalloc(&ctx->d_c_alt, (size_t)bns->n_seqs));
		CK(cudaMemcpy(ctx->d_c_alt, bns->is_alt, (size_t)bns->n_seqs, cudaMemcpyHostToDevice));
	}
	ctx->l_pac = bns->l_pac; ctx->n_seqs = bns->n_seqs; ctx->copt = *opt;
	// most reads need one B-tree node (up to nine chains); the rest a third of their seeds (k_chain_node_counts)
	ctx->node_cap = 2 * (uint64_t)ctx->max_reads + ctx->max_seeds / 4 + 4096;   // see k_chain_node_counts
	for (int i = 0; i < ctx->n_slots; ++i) {
		Slot *s = &ctx->slots[i];
		if (s->d_s_next) continue;
		const size_t ns = ctx->max_seeds, nr = (size_t)ctx->max_reads + 1;
		CK(cudaMalloc(&s->d_s_next, ns * 4)); CK(cudaMalloc(&s->d_s_qb_len, ns * 4)); CK(cudaMalloc(&s->d_order, ns * 4)); CK(cudaMalloc(&s->d_klist, ns * 4));
		CK(cudaMalloc(&s->d_chain_tmp, ns * sizeof(ChainTmp)));
		CK(cudaMalloc(&s->d_node_cnt, nr * 4)); CK(cudaMalloc(&s->d_node_off, nr * 4));
		CK(cudaMalloc(&s->d_nodes, ctx->node_cap * 24 * 4));
		CK(cudaMalloc(&s->d_n_chain, nr * 4)); CK(cudaMalloc(&s->d_n_cseed, nr * 4)); CK(cudaMalloc(&s->d_l_rep, nr * 4));
		CK(cudaMalloc(&s->d_chain_off, nr * 4)); CK(cudaMalloc(&s->d_cseed_off, nr * 4));
		CK(cudaMalloc(&s->d_chains, ctx->max_mems * sizeof(cs_chain_t)));
		CK(cudaMalloc(&s->d_cs_lo, ns * 4)); CK(cudaMalloc(&s->d_cs_hi, ns)); CK(cudaMalloc(&s->d_cs_qbeg, ns * 2)); CK(cudaMalloc(&s->d_cs_len, ns * 2));
	}
	ctx->chaining = true;
	return CS_OK;
fail:
	ctx->chaining = false;
	return CS_E_CUDA;
}

extern "C" int cs_seed_batch_wait_chains(cs_ctx_t *ctx, int slot, cs_chain_result_t *out)
{
	int rc;
	if ((rc = check_slot(ctx, slot)) != CS_OK) return rc;
	if (!out) return set_err(CS_E_ARG, "null result");
	Slot *s = &ctx->slots[slot];
	if (s->state != 2 && s->state != 3) return set_err(CS_E_STATE, "slot %d has no batch in flight or finished", slot);
	if (!s->chained) return set_err(CS_E_STATE, "the batch in slot %d was not chained (cs_ctx_set_chaining)", slot);
	{ const int rc_ = use_device(ctx->device); if (rc_ != CS_OK) return rc_; }
	if (s->state == 2 && (rc = finish_run(ctx, s)) != CS_OK) { s->state = 1; return rc; }
	{
		const uint32_t n = s->n_reads;
		const Ctrl *h = s->h_ctrl;
		if (h->n_chains > ctx->max_mems || h->n_cseeds > ctx->max_seeds)
			return set_err(CS_E_OVERFLOW, "chain buffers too small: %u chains of %llu, %u chain seeds of %llu", h->n_chains, (unsigned long long)ctx->max_mems, h->n_cseeds, (unsigned long long)ctx->max_seeds);
		if (!s->h_chain_off) {
			CK(cudaMallocHost(&s->h_chain_off, ((size_t)ctx->max_reads + 1) * 4)); CK(cudaMallocHost(&s->h_cseed_off, ((size_t)ctx->max_reads + 1) * 4));
			CK(cudaMallocHost(&s->h_chains, ctx->max_mems * sizeof(cs_chain_t)));
			CK(cudaMallocHost(&s->h_cs_lo, ctx->max_seeds * 4)); CK(cudaMallocHost(&s->h_cs_hi, ctx->max_seeds));
			CK(cudaMallocHost(&s->h_cs_qbeg, ctx->max_seeds * 2)); CK(cudaMallocHost(&s->h_cs_len, ctx->max_seeds * 2));
		}
		CK(cudaMemcpyAsync(s->h_chain_off, s->d_chain_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
		CK(cudaMemcpyAsync(s->h_cseed_off, s->d_cseed_off, ((size_t)n + 1) * 4, cudaMemcpyDeviceToHost, s->stream_out));
		if (h->n_chains) CK(cudaMemcpyAsync(s->h_chains, s->d_chains, (size_t)h->n_chains * sizeof(cs_chain_t), cudaMemcpyDeviceToHost, s->stream_out));
		if (h->n_cseeds) {
			CK(cudaMemcpyAsync(s->h_cs_lo, s->d_cs_lo, (size_t)h->n_cseeds * 4, cudaMemcpyDeviceToHost, s->stream_out));
			CK(cudaMemcpyAsync(s->h_cs_hi, s->d_cs_hi, (size_t)h->n_cseeds, cudaMemcpyDeviceToHost, s->stream_out));
			CK(cudaMemcpyAsync(s->h_cs_qbeg, s->d_cs_qbeg, (size_t)h->n_cseeds * 2, cudaMemcpyDeviceToHost, s->stream_out));
			CK(cudaMemcpyAsync(s->h_cs_len, s->d_cs_len, (size_t)h->n_cseeds * 2, cudaMemcpyDeviceToHost, s->stream_out));
		}
		CK(cudaEventRecord(s->ev_done, s->stream_out));
		CK(cudaEventSynchronize(s->ev_done));
		s->state = 3;
		memset(out, 0, sizeof *out);
		out->n_reads = n; out->n_chains = h->n_chains; out->n_cseeds = h->n_cseeds;
		out->chain_off = s->h_chain_off; out->cseed_off = s->h_cseed_off; out->chains = s->h_chains;
		out->rbeg_lo = s->h_cs_lo; out->rbeg_hi = s->h_cs_hi; out->qbeg = s->h_cs_qbeg; out->len = s->h_cs_len;
	}
	return CS_OK;
fail:
	return CS_E_CUDA;
}

extern "C" int cs_compact_expand(const cs_compact_result_t *res, cs_mem_t *mems, int64_t *rbeg, int n_threads)
{ // host-side C++, no device involved
	if (!res || (res->n_mems && !mems) || (res->n_seeds && !rbeg)) return set_err(CS_E_ARG, "null argument");
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 64) n_threads = 64;
	auto work = [=](int t) {
		const uint64_t m0 = res->n_mems * t / n_threads, m1 = res->n_mems * (t + 1) / n_threads;
		for (uint64_t i = m0; i < m1; ++i) cs_cmem_unpack(res->cmems + i, mems + i);
		const uint64_t s0 = res->n_seeds * t / n_threads, s1 = res->n_seeds * (t + 1) / n_threads;
		for (uint64_t i = s0; i < s1; ++i) rbeg[i] = cs_crbeg(res->rbeg_lo, res->rbeg_hi, i);
	};
	if (n_threads == 1) { work(0); return CS_OK; }
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
	for (auto &t : th) t.join();
	return CS_OK;
}

// ---------------------------------------------------------------------------------------------
// measurement helpers
// ---------------------------------------------------------------------------------------------
extern "C" int cs_probe_random_gather_ex(int device, uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iters, int unroll,
                                         int l2_fetch_granularity, double *gbytes_per_s, double *gloads_per_s);

extern "C" int cs_probe_random_gather(int device, uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iters,
                                      double *gbytes_per_s, double *gloads_per_s)
{
	return cs_probe_random_gather_ex(device, table_bytes, granule, n_loads, iters, 1, 0, gbytes_per_s, gloads_per_s);
}

extern "C" int cs_probe_random_gather_ex(int device, uint64_t table_bytes, uint32_t granule, uint64_t n_loads, int iters, int unroll,
                                         int l2_fetch_granularity, double *gbytes_per_s, double *gloads_per_s)
{
	uint4 *d_t = nullptr; unsigned long long *d_sink = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	float best = 1e30f;
	if (granule != 16 && granule != 32 && granule != 64) return set_err(CS_E_ARG, "granule must be 16, 32 or 64");
	{ const int rc_ = use_device(device); if (rc_ != CS_OK) return rc_; }
	if (l2_fetch_granularity) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)l2_fetch_granularity);
	{
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		uint64_t n16 = table_bytes / 16;
		CK(cudaMalloc(&d_t, n16 * 16)); CK(cudaMalloc(&d_sink, 8));
		CK(cudaMemset(d_sink, 0, 8));
		k_fill<<<prop.multiProcessorCount * 8, 256>>>(d_t, n16, 7u);
		CK(cudaGetLastError());
		CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
		for (int it = 0; it < iters + 1; ++it) {
			CK(cudaEventRecord(e0));
			if (unroll >= 4) k_gather_probe4<<<prop.multiProcessorCount * 8, 256>>>(d_t, table_bytes / granule, granule / 16, n_loads, 0x1234567ull + it, d_sink);
			else k_gather_probe<<<prop.multiProcessorCount * 8, 256>>>(d_t, table_bytes / granule, granule / 16, n_loads, 0x1234567ull + it, d_sink);
			CK(cudaGetLastError());
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			float ms;
			CK(cudaEventElapsedTime(&ms, e0, e1));
			if (it > 0 && ms < best) best = ms;
		}
	}
	if (gloads_per_s) *gloads_per_s = (double)n_loads / (best * 1e-3) * 1e-9;
	if (gbytes_per_s) *gbytes_per_s = (double)n_loads * granule / (best * 1e-3) * 1e-9;
	cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_t); cudaFree(d_sink);
	return CS_OK;
fail:
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	cudaFree(d_t); cudaFree(d_sink);
	return CS_E_CUDA;
}

extern "C" int cs_debug_stats(cs_ctx_t *ctx, int slot, uint64_t out[40])
{
	if (check_slot(ctx, slot) != CS_OK) return CS_E_ARG;
	if (!out) return set_err(CS_E_ARG, "null argument");
	for (int k = 0; k < 20; ++k) out[k] = ctx->slots[slot].h_ctrl->counters[k];
	out[20] = ctx->slots[slot].h_ctrl->n_defer;
	out[21] = ctx->grid_fast; out[22] = ctx->grid; out[23] = ctx->slots[slot].h_ctrl->n_lit;
	for (int k = 0; k < 16; ++k) out[24 + k] = ctx->slots[slot].h_ctrl->counters[20 + k];
	return CS_OK;
}

extern "C" int cs_host_register(void *ptr, size_t bytes)
{
	if (cs_device_count() <= 0) return set_err(CS_E_NODEVICE, "no CUDA device visible: compseed_b200 has no CPU path");
	cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) { cudaGetLastError(); return set_err(CS_E_CUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
	return CS_OK;
}

extern "C" int cs_host_unregister(void *ptr)
{
	cudaError_t e = cudaHostUnregister(ptr);
	if (e != cudaSuccess) { cudaGetLastError(); return set_err(CS_E_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
	return CS_OK;
}

extern "C" int cs_flush_l2(int device)
{
	static uint4 *buf[16] = {nullptr};
	const uint64_t bytes = 512ull << 20;
	{ const int rc_ = use_device(device); if (rc_ != CS_OK) return rc_; }
	if (device >= 16) return set_err(CS_E_ARG, "device index too large");
	if (!buf[device]) CK(cudaMalloc(&buf[device], bytes));
	k_fill<<<1024, 256>>>(buf[device], bytes / 16, 3u);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	return CS_OK;
fail:
	return CS_E_CUDA;
}
