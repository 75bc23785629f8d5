// emu_gf.cpp -- runs gapped_filter_kernel (diamond_b200/csrc/cuda/gf_kernels.cuh, the source the GPU library is built from) on
// the CPU behind tests/emu_cuda.h and checks it against the oracle's dmnd_hits_gapped_filter on the hits of a real seed search
// (--sensitive parameters, Hauser bias on, SEED_MASK bits still set on the query letters as in the pipeline).
// usage: emu_gf DIR MAX_HITS      (DIR holds q.i8 q.i64 r.i8 r.i64: the two block images)
#include "emu_cuda.h"
#include "../diamond_b200/csrc/cuda/gf_kernels.cuh"
#include <string>
using namespace dmnd_cuda;

template<typename T> static std::vector<T> slurp(const std::string& path) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
	fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<T> v((size_t)n / sizeof(T));
	if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) exit(2);
	fclose(f);
	return v;
}

int main(int argc, char** argv) {
	if (argc < 3) return 2;
	const std::string dir = argv[1];
	const size_t max_hits = (size_t)atol(argv[2]);
	std::vector<int8_t> qraw = slurp<int8_t>(dir + "/q.i8"), rraw = slurp<int8_t>(dir + "/r.i8");
	std::vector<int64_t> qlim = slurp<int64_t>(dir + "/q.i64"), rlim = slurp<int64_t>(dir + "/r.i64");
	const uint32_t nq = (uint32_t)qlim.size() - 1, nr = (uint32_t)rlim.size() - 1;
	dmnd_search_opts o; dmnd_search_opts_default(&o); o.sensitivity = 3;
	if (argc > 3) o.query_contexts = atoi(argv[3]);  // 6: translated frames (cutoffs at the first frame's length, one scan for short queries)
	dmnd_params hp; if (dmnd_params_init(&o, &hp)) { fprintf(stderr, "%s\n", dmnd_last_error()); return 2; }
	dmnd_ctx* ctx; if (dmnd_create(0, &hp, &ctx)) return 2;
	dmnd_block *qb, *rb;
	if (dmnd_block_upload(ctx, qraw.data(), qraw.size(), qlim.data(), nq, &qb) || dmnd_block_upload(ctx, rraw.data(), rraw.size(), rlim.data(), nr, &rb)) return 2;
	if (dmnd_block_compute_bias(ctx, qb, 1)) return 2;
	int fails = 0; size_t total = 0, passed = 0;
	for (int sid = 0; sid < 2; ++sid) {
		dmnd_hits* h; dmnd_stage_counters cn;
		if (dmnd_search_shape(ctx, qb, rb, sid, &h, &cn)) { fprintf(stderr, "%s\n", dmnd_last_error()); return 2; }
		size_t n = dmnd_hits_count(h);
		std::vector<dmnd_hit> hits(n);
		if (n) dmnd_hits_download(ctx, h, hits.data(), n);
		std::vector<uint8_t> want(n, 9);
		if (n && dmnd_hits_gapped_filter(ctx, qb, rb, h, want.data(), n)) { fprintf(stderr, "%s\n", dmnd_last_error()); return 2; }
		dmnd_hits_free(ctx, h);
		if (n > max_hits) n = max_hits;
		// what the device kernel sees: the block images with SEED_MASK bits and the bias array
		std::vector<int8_t> ql(qraw.size()), bias(qraw.size());
		dmnd_block_download_letters(ctx, qb, ql.data(), ql.size());
		dmnd_block_download_bias(ctx, qb, bias.data(), bias.size());
		DevParams P; memset(&P, 0, sizeof P);
		memcpy(P.score, hp.score, 1024); P.gap_open = hp.gap_open; P.gap_extend = hp.gap_extend;
		memcpy(P.gapped_cutoff1, hp.gapped_cutoff1, sizeof P.gapped_cutoff1); memcpy(P.gapped_cutoff2, hp.gapped_cutoff2, sizeof P.gapped_cutoff2);
		P.gapped_filter_diag_score = hp.gapped_filter_diag_score; P.gapped_filter_window = hp.gapped_filter_window;
		P.query_contexts = hp.query_contexts > 1 ? hp.query_contexts : 1;
		std::vector<uint8_t> got(n + 8, 7);
		emu::launch((unsigned)((n + 3) / 4), 128, [&] { gapped_filter_kernel(ql.data(), bias.data(), qlim.data(), rraw.data(), rlim.data(), nr, hits.data(), n, &P, got.data()); });
		size_t bad = 0;
		for (size_t k = 0; k < n; ++k) { bad += got[k] != want[k]; passed += want[k]; }
		total += n;
		if (bad) { ++fails; printf("FAIL shape %d: %zu of %zu pass flags differ\n", sid, bad, n); }
	}
	printf("hits=%zu pass=%zu fails=%d \n", total, passed, fails);
	return fails ? 1 : 0;
}
