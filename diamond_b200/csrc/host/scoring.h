// scoring.h -- BLOSUM62 11/1 scoring system, e-value / bit-score statistics and the Hauser composition bias,
// restated on the host (double / float arithmetic identical to the reference's scalar code).
//   stats/score_matrix.cpp:43-70,217-254   ScoreMatrix ctor, evalue, bitscore, rawscore
//   lib/alp/sls_pvalues.cpp:367-507        finite-size corrected area (get_appr_tail_prob_with_cov_without_errors)
//   lib/alp/sls_alignment_evaluer.cpp:656-841, .hpp:135-161   parameter mapping, evalue = area * K * exp(-lambda*s)
//   stats/hauser_correction.cpp:53-109     HauserCorrection
#pragma once
#include <cstdint>
#include <vector>
#include "../../../include/dmnd_b200.h"

namespace dmnd {

constexpr int ALPHABET = 26;  // ARNDCQEGHILKMFPSTWYVBJZX*_  (basic/value.h:53)
constexpr int MASK_LETTER = 23, STOP_LETTER = 24, SUPER_HARD_MASK = 25;

struct Scoring {
	int8_t m8[32 * 32];
	int m32[32 * 32];
	int gap_open = 11, gap_extend = 1;
	double lambda, K, ln_k;
	// ALP parameters (Sls::AlignmentEvaluerParameters as built by alp_params(), score_matrix.cpp:43-47)
	double a_I, b_I, alpha_I, beta_I, a_J, b_J, alpha_J, beta_J, sigma, tau;
	double vi_y_thr, vj_y_thr, c_y_thr;
	double db_letters = 0;
	double background_scores[20];
	int raw_ungapped_xdrop;

	Scoring();
	int score(int a, int b) const { return m32[a * 32 + b]; }
	double area(double y, double m, double n) const;
	double evalue(int raw_score, unsigned qlen, unsigned slen) const;
	double bitscore(double raw_score) const;
	int rawscore(double bits) const;
};

// int8 per-position bias, padded with 32 zeros (HauserCorrection::int8).
void hauser_correction(const Scoring& sc, const int8_t* seq, int len, std::vector<int8_t>& out);

const char* letter_alphabet();

}  // namespace dmnd
