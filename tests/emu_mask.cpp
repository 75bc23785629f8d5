// emu_mask.cpp -- runs the masking kernels of diamond_b200/csrc/cuda/mask_kernels.cuh (the SOURCE the GPU library is built
// from) on the CPU behind tests/emu_cuda.h and checks them against the oracle: tantan letters + positions, the motif
// soft-masking table, the SEED_MASK marking of the query side.  The launch sequence mirrors block_mask_impl (mask.cu) with
// the same grid and block sizes; whole-block calls and per-range calls (what the query lanes do) must agree.
// usage: emu_mask DIR NSEQ_LIMIT     (DIR holds raw.i8 and lim.i64 of one block image)
#include "emu_cuda.h"
#include "../diamond_b200/csrc/cuda/mask_kernels.cuh"
#include <algorithm>
#include <string>
using namespace dmnd_cuda;

extern "C" int dmnd_oracle_block_soft(const dmnd_block* b, uint8_t* out, size_t raw_len);
extern "C" int dmnd_oracle_motif_seed_mask(dmnd_ctx* ctx, dmnd_block* b, int sid, uint32_t q_begin, uint32_t q_end);

template<typename T> static std::vector<T> slurp(const std::string& path) {
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
	fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<T> v((size_t)n / sizeof(T));
	if (fread(v.data(), 1, (size_t)n, f) != (size_t)n) exit(2);
	fclose(f);
	return v;
}

static size_t g_forward_seqs = 0, g_backward_seqs = 0;
struct EmuBlock {
	std::vector<int8_t> letters;  // with the slack the device allocation has
	std::vector<int64_t> limits;
	std::vector<uint32_t> soft;
	size_t raw_len; uint32_t nseq;
};

// mirrors block_mask_impl (diamond_b200/csrc/cuda/mask.cu)
static void emu_block_mask(const DevParams& P, EmuBlock& b, int algo, uint32_t s_begin, uint32_t s_end, unsigned tantan_block, std::vector<uint64_t>& pos) {
	pos.clear();
	const size_t p_begin = (size_t)b.limits[s_begin], p_end = (size_t)b.limits[s_end];
	const size_t nlet = p_end - p_begin, nseq = s_end - s_begin;
	if (nseq == 0 || nlet == 0) return;
	const size_t w_begin = p_begin >> 5, w_end = ((p_end - 1) >> 5) + 1;
	std::vector<uint32_t> bits((b.raw_len >> 5) + 4, 0xdeadbeefu);  // scratch is NOT zero on the device either
	if (algo & DMND_MASK_TANTAN) {
		std::vector<float> pb(nlet, -1.0f), scale((nlet >> 4) + nseq + 2, -1.0f);
		std::fill(bits.begin() + (ptrdiff_t)w_begin, bits.begin() + (ptrdiff_t)w_end, 0u);
		std::vector<uint32_t> len(nseq), id(nseq);
		emu::launch((unsigned)((nseq + 255) / 256), 256, [&] { seq_len_kernel(b.limits.data(), s_begin, (uint32_t)nseq, len.data(), id.data()); });
		std::vector<uint32_t> order(id);  // cub::DeviceRadixSort::SortPairsDescending (stable) on (len, id)
		std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return len[x - s_begin] > len[y - s_begin]; });
		std::vector<float> zinv(nseq, -1.0f);
		std::vector<uint8_t> need(nseq + 16, 0xee);
		emu::launch((unsigned)((nseq + 15) / 16), tantan_block, [&] { tantan_forward_kernel(b.letters.data(), b.limits.data(), order.data(), (uint32_t)nseq, &P, pb.data(), scale.data(),
			(int64_t)p_begin, s_begin, zinv.data(), need.data()); });
		std::vector<uint32_t> order2;  // cub::DeviceSelect::Flagged(order, need)
		for (size_t x = 0; x < nseq; ++x) { if (need[x] > 1) { printf("FAIL need flag of slot %zu not written\n", x); exit(1); } if (need[x]) order2.push_back(order[x]); }
		int n_need = (int)order2.size();
		g_backward_seqs += (size_t)n_need; g_forward_seqs += nseq;
		order2.resize(nseq + 1, 0xffffffffu);
		emu::launch((unsigned)((nseq + 15) / 16), tantan_block, [&] { tantan_backward_kernel(b.letters.data(), b.limits.data(), order2.data(), &n_need, &P, pb.data(), scale.data(),
			(int64_t)p_begin, s_begin, zinv.data(), bits.data()); });
		unsigned long long total = 0;
		emu::launch((unsigned)((w_end - w_begin + 255) / 256), 256, [&] { popc_kernel(bits.data(), w_begin, w_end, &total); });
		BitSet sel{ bits.data() };
		for (uint64_t p = p_begin; p < p_end; ++p) if (sel(p)) pos.push_back(p);  // cub::DeviceSelect::If over a counting sequence
		if (pos.size() != total) { printf("FAIL popc_kernel total %llu != selected %zu\n", total, pos.size()); exit(1); }
	}
	if (algo & DMND_MASK_MOTIF) {
		std::vector<uint32_t> flag((size_t)b.nseq / 32 + 2, 0), seqs(nseq + 1, 0xffffffffu);
		std::vector<uint64_t> table(DMND_MOTIF_CODES, DMND_MOTIF_CODES + DMND_MOTIF_COUNT);
		unsigned int nlist = 0;
		std::fill(bits.begin() + (ptrdiff_t)w_begin, bits.begin() + (ptrdiff_t)w_end + 1, 0u);
		emu::launch((unsigned)((w_end - w_begin + 255) / 256), 256, [&] { clear_bits_kernel(b.soft.data(), p_begin, p_end); });
		emu::launch((unsigned)((nlet + MOTIF_TILE - 1) / MOTIF_TILE), 256, [&] {
			motif_hit_kernel(b.letters.data(), b.raw_len, p_begin, p_end, table.data(), bits.data(), b.limits.data(), s_begin, s_end, flag.data(), seqs.data(), &nlist); });
		emu::launch((unsigned)((nseq + 127) / 128), 128, [&] { motif_apply_kernel(b.limits.data(), seqs.data(), &nlist, bits.data(), b.soft.data(), P.max_motif_len); });
	}
}

int main(int argc, char** argv) {
	if (argc < 3) return 2;
	const std::string dir = argv[1];
	const uint32_t want = (uint32_t)atoi(argv[2]);
	std::vector<int8_t> raw = slurp<int8_t>(dir + "/raw.i8");
	std::vector<int64_t> lim = slurp<int64_t>(dir + "/lim.i64");
	uint32_t nseq = (uint32_t)lim.size() - 1;
	if (want && want < nseq) {  // truncate to the first `want` sequences (re-pad the tail)
		nseq = want; lim.resize(nseq + 1);
		raw.resize((size_t)lim[nseq]); raw.insert(raw.end(), DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER);
	}
	dmnd_search_opts o; dmnd_search_opts_default(&o);
	dmnd_params hp; if (dmnd_params_init(&o, &hp)) return 2;
	DevParams P; memset(&P, 0, sizeof P);
	memcpy(P.tantan_lr, hp.tantan_lr, sizeof P.tantan_lr); memcpy(P.tantan_d, hp.tantan_d, sizeof P.tantan_d);
	P.tantan_b2b = hp.tantan_b2b; P.tantan_f2f = hp.tantan_f2f; P.tantan_p_repeat_end = hp.tantan_p_repeat_end; P.tantan_p_mask = hp.tantan_p_mask;
	P.max_motif_len = hp.max_motif_len;
	memcpy(P.shape_len, hp.shape_len, sizeof P.shape_len);
	// ---- oracle
	dmnd_ctx* ctx; if (dmnd_create(0, &hp, &ctx)) return 2;
	dmnd_block* ob; if (dmnd_block_upload(ctx, raw.data(), raw.size(), lim.data(), nseq, &ob)) return 2;
	uint64_t n_hard = 0;
	if (dmnd_block_mask(ctx, ob, DMND_MASK_TANTAN | DMND_MASK_MOTIF, 0, nseq, &n_hard)) return 2;
	std::vector<uint64_t> opos((size_t)n_hard);
	dmnd_block_mask_fetch(ctx, opos.data(), opos.size());
	std::vector<int8_t> olet(raw.size()); dmnd_block_download_letters(ctx, ob, olet.data(), olet.size());
	std::vector<uint8_t> osoft(raw.size()); dmnd_oracle_block_soft(ob, osoft.data(), osoft.size());
	dmnd_oracle_motif_seed_mask(ctx, ob, 0, 0, nseq);
	std::vector<int8_t> oseed(raw.size()); dmnd_block_download_letters(ctx, ob, oseed.data(), oseed.size());
	size_t n_soft = 0, n_bits = 0;
	for (uint8_t x : osoft) n_soft += x;
	for (size_t i = 0; i < oseed.size(); ++i) n_bits += oseed[i] != DMND_DELIMITER && (oseed[i] & 0x80);
	int fails = 0;
	// ---- emulated kernels: (a) whole block with the library's block size, (b) three ranges with a small block
	for (int mode = 0; mode < 2; ++mode) {
		EmuBlock b;
		b.letters = raw; b.letters.resize(((raw.size() + 63) & ~(size_t)63) + 64, (int8_t)DMND_DELIMITER);
		b.limits = lim; b.raw_len = raw.size(); b.nseq = nseq;
		b.soft.assign(b.letters.size() / 32 + 4, mode == 0 ? 0u : 0xffffffffu);  // (b): stale bits from an earlier call must be cleared
		if (mode == 1) {  // ... but only inside the sequence area: the padding on either side is zero since the upload
			for (size_t w = 0; w < 8; ++w) b.soft[w] = 0;
			for (size_t p = (size_t)lim[nseq]; p < b.soft.size() * 32; ++p) b.soft[p >> 5] &= ~(1u << (p & 31));
		}
		std::vector<uint64_t> pos, all;
		std::vector<uint32_t> cuts = mode == 0 ? std::vector<uint32_t>{ 0, nseq } : std::vector<uint32_t>{ 0, nseq / 3, nseq / 3, (2 * nseq) / 3 + 1, nseq };
		for (size_t c = 0; c + 1 < cuts.size(); ++c) {
			emu_block_mask(P, b, DMND_MASK_TANTAN | DMND_MASK_MOTIF, cuts[c], cuts[c + 1], 128u, pos);
			all.insert(all.end(), pos.begin(), pos.end());
		}
		if (all != opos) { ++fails; printf("FAIL mode %d: %zu masked positions, oracle %zu\n", mode, all.size(), opos.size()); }
		if (memcmp(b.letters.data(), olet.data(), raw.size()) != 0) { ++fails; printf("FAIL mode %d: letters after tantan differ\n", mode); }
		size_t soft_diff = 0;
		for (size_t p = 0; p < raw.size(); ++p) soft_diff += ((b.soft[p >> 5] >> (p & 31)) & 1u) != osoft[p];
		if (soft_diff) { ++fails; printf("FAIL mode %d: %zu soft bits differ\n", mode, soft_diff); }
		// SEED_MASK marking of the query side (motif_seedmask_kernel in dmnd_search_shape), per range as well
		for (size_t c = 0; c + 1 < cuts.size(); ++c) {
			const size_t pb = (size_t)lim[cuts[c]], pe = (size_t)lim[cuts[c + 1]];
			if (pe > pb) emu::launch((unsigned)((pe - pb + 255) / 256), 256, [&] { motif_seedmask_kernel(b.letters.data(), b.soft.data(), pb, pe, P.shape_len[0]); });
		}
		if (memcmp(b.letters.data(), oseed.data(), raw.size()) != 0) { ++fails; printf("FAIL mode %d: SEED_MASK marking differs\n", mode); }
	}
	printf("seqs=%u letters=%zu tantan_masked=%zu soft_letters=%zu seed_mask_positions=%zu context_switches=%llu backward_pass_seqs=%zu/%zu fails=%d \n", nseq, raw.size() - 512 - nseq, opos.size(),
	       n_soft, n_bits, (unsigned long long)emu::g_switches, g_backward_seqs, g_forward_seqs, fails);
	return fails ? 1 : 0;
}
