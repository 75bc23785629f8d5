// dev_params.h -- device-side copy of what the kernels need from dmnd_params.  Plain C++ (no CUDA headers) so that the
// CPU emulation of the kernels under tests/ can include the kernel sources unchanged.
#pragma once
#include <cstdint>
#include "../../../include/dmnd_b200.h"

namespace dmnd_cuda {

// Device-side copy of what the kernels need from dmnd_params (constant-memory sized).
struct DevParams {
	int8_t score[1024];
	uint8_t reduction[32], map8[32], map8b[32];
	int32_t shape_pos[DMND_MAX_SHAPES][DMND_MAX_WEIGHT];
	uint32_t shape_mask[DMND_MAX_SHAPES];
	int32_t shape_len[DMND_MAX_SHAPES];
	int32_t n_shapes, shape_weight, reduction_size;
	int32_t hamming_id, seedp_bits, index_chunks, left_most_interval, ungapped_window;
	int32_t gap_open, gap_extend;
	int32_t seed_bits;     // bit length of reduction_size^weight - 1
	double seed_cut;
	double lnfact[DMND_MAX_WEIGHT + 1];
	float background_scores_f32[20];
	int32_t ungapped_cutoff[32], short_query_ungapped_cutoff, short_query_max_len;  // stage-2 ungapped window filter (0 tables: skipped)
	int32_t query_contexts, ungapped_cutoff_short[32];  // blastx: 6 frames per query; frames of <= 85 letters (search/stage2.h:41-63)
	int16_t gapped_cutoff1[32][32], gapped_cutoff2[32][32];  // gapped filter (align/gapped_filter.cpp): CutoffTable2D at e-values 2000 / gapped_filter_evalue
	int32_t gapped_filter_diag_score, gapped_filter_window;
	// tantan (masking/tantan.cpp:121-214): likelihood ratios [a*32+b], per-offset repeat start probabilities, transition constants
	float tantan_lr[1024];
	float tantan_d[50];
	float tantan_b2b, tantan_f2f, tantan_p_repeat_end, tantan_p_mask;
	int32_t max_motif_len;
	uint32_t one, k65536;  // 1 and 65536 as run-time values: products with them are IMADs (FMA pipe) where a constant would become an ALU-pipe add / shift / PRMT (swipe16.cuh)
	uint32_t neg2;  // 0x80008000 = (-32768, -32768): third operand of the packed bias add of swipe16.cuh (max with it is the identity)
	uint32_t zero;  // always 0: read into a register where ptxas would otherwise materialise a packed 0 with an extra PRMT per use (swipe16.cuh)
};

}  // namespace dmnd_cuda
