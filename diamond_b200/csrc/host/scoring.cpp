#include "scoring.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace dmnd {

const char* letter_alphabet() { return "ARNDCQEGHILKMFPSTWYVBJZX*_"; }

// BLOSUM62 in the reference's letter order ARNDCQEGHILKMFPSTWYV BJZX* (NCBI BLOSUM62; X row -1, * row -4, */* = 1).
// Column '_' (SUPER_HARD_MASK) scores -12 against everything (stats/matrices/blosum62.h:97-98).
static const int8_t B62[25][25] = {
	/*A*/ { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1,-1,-1,-4},
	/*R*/ {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1,-2, 0,-1,-4},
	/*N*/ {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 4,-3, 0,-1,-4},
	/*D*/ {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4,-3, 1,-1,-4},
	/*C*/ { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-1,-3,-1,-4},
	/*Q*/ {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0,-2, 4,-1,-4},
	/*E*/ {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1,-3, 4,-1,-4},
	/*G*/ { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-4,-2,-1,-4},
	/*H*/ {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0,-3, 0,-1,-4},
	/*I*/ {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3, 3,-3,-1,-4},
	/*L*/ {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4, 3,-3,-1,-4},
	/*K*/ {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0,-3, 1,-1,-4},
	/*M*/ {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3, 2,-1,-1,-4},
	/*F*/ {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3, 0,-3,-1,-4},
	/*P*/ {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-3,-1,-1,-4},
	/*S*/ { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0,-2, 0,-1,-4},
	/*T*/ { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1,-1,-1,-4},
	/*W*/ {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-2,-2,-1,-4},
	/*Y*/ {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-1,-2,-1,-4},
	/*V*/ { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3, 2,-2,-1,-4},
	/*B*/ {-2,-1, 4, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4,-3, 0,-1,-4},
	/*J*/ {-1,-2,-3,-3,-1,-2,-3,-4,-3, 3, 3,-3, 2, 0,-3,-2,-1,-2,-1, 2,-3, 3,-3,-1,-4},
	/*Z*/ {-1, 0, 0, 1,-3, 4, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-2,-2,-2, 0,-3, 4,-1,-4},
	/*X*/ {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-4},
	/***/ {-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4, 1},
};

// Robinson-Robinson background frequencies used by the reference for BLOSUM62 (stats/matrices/blosum62.h:240-246).
static const double BG[20] = {
	7.4216205067993410e-02, 5.1614486141284638e-02, 4.4645808512757915e-02, 5.3626000838554413e-02,
	2.4687457167944848e-02, 3.4259650591416023e-02, 5.4311925684587502e-02, 7.4146941452644999e-02,
	2.6212984805266227e-02, 6.7917367618953756e-02, 9.8907868497150955e-02, 5.8155682303079680e-02,
	2.4990197579643110e-02, 4.7418459742284751e-02, 3.8538003320306206e-02, 5.7229029476494421e-02,
	5.0891364550287033e-02, 1.3029956129972148e-02, 3.2281512313758580e-02, 7.2919098205619245e-02 };

Scoring::Scoring() {
	for (int i = 0; i < 32; ++i)
		for (int j = 0; j < 32; ++j) {
			int v;
			if (i < 25 && j < 25) v = B62[i][j];
			else if (i < ALPHABET && j < ALPHABET) v = -12;
			else v = -128;  // SCHAR_MIN outside the alphabet (stats/score_matrix.cpp: Scores ctor)
			m32[i * 32 + j] = v;
			m8[i * 32 + j] = (int8_t)v;
		}
	// Gapped constants for BLOSUM62 11/1 and the ungapped row (stats/matrices/blosum62.h:29,38).
	const double Lambda = 0.267, Kp = 0.041, alpha = 1.9, alpha_v = 42.6028, sig = 43.6362;
	const double u_alpha = 0.7916, u_alpha_v = 4.964660;
	const double G = 11 + 1;
	lambda = Lambda; K = Kp; ln_k = std::log(K);
	const double b = 2.0 * G * (u_alpha - alpha), beta = 2.0 * G * (u_alpha_v - alpha_v);
	a_I = alpha; b_I = b; a_J = alpha; b_J = b;
	alpha_I = alpha_v; beta_I = beta; alpha_J = alpha_v; beta_J = beta;
	sigma = sig; tau = 2.0 * G * (u_alpha_v - sig);
	// pvalues::compute_tmp_values (sls_pvalues.cpp:343-362), nat_cut_off_in_max = 2
	vi_y_thr = std::max(2.0 * alpha_I / lambda, 0.0);
	vj_y_thr = std::max(2.0 * alpha_J / lambda, 0.0);
	c_y_thr = std::max(2.0 * sigma / lambda, 0.0);
	for (int i = 0; i < 20; ++i) {
		background_scores[i] = 0;
		for (int j = 0; j < 20; ++j) background_scores[i] += BG[j] * score(i, j);
	}
	raw_ungapped_xdrop = rawscore(12.3);  // basic/config.cpp:428,853
}

static inline double normal_probability(double x) { return 0.5 * std::erfc(-0.70710678118654752440 * x); }

double Scoring::area(double y, double m, double n) const {
	const double const_val = 1.0 / std::sqrt(2.0 * 3.1415926535897932384626433832795);
	const double m_li_y = m - (a_I * y + b_I);
	const double vi_y = std::max(vi_y_thr, alpha_I * y + beta_I);
	const double sqrt_vi_y = std::sqrt(vi_y);
	const double m_F = sqrt_vi_y == 0.0 ? 1e100 : m_li_y / sqrt_vi_y;
	const double P_m_F = normal_probability(m_F);
	const double E_m_F = -const_val * std::exp(-0.5 * m_F * m_F);
	const double p1 = m_li_y * P_m_F - sqrt_vi_y * E_m_F;
	const double n_lj_y = n - (a_J * y + b_J);
	const double vj_y = std::max(vj_y_thr, alpha_J * y + beta_J);
	const double sqrt_vj_y = std::sqrt(vj_y);
	const double n_F = sqrt_vj_y == 0.0 ? 1e100 : n_lj_y / sqrt_vj_y;
	const double P_n_F = normal_probability(n_F);
	const double E_n_F = -const_val * std::exp(-0.5 * n_F * n_F);
	const double p2 = n_lj_y * P_n_F - sqrt_vj_y * E_n_F;
	const double c_y = std::max(c_y_thr, sigma * y + tau);
	return p1 * p2 + c_y * P_m_F * P_n_F;
}

double Scoring::evalue(int raw_score, unsigned qlen, unsigned slen) const {
	// AlignmentEvaluer::evalue(score, seqlen1 = qlen, seqlen2 = slen): area(score, qlen, slen) passes (m_ = seqlen2, n_ = seqlen1)
	const double s = (double)raw_score;
	return area(s, (double)slen, (double)qlen) * (K * std::exp(-lambda * s)) * db_letters / (double)slen;
}

double Scoring::bitscore(double raw_score) const {
	const double s = std::round(raw_score);
	return (lambda * s - ln_k) / 0.69314718055994530941723212145818;
}

int Scoring::rawscore(double bits) const {
	return (int)std::ceil((bits * 0.69314718055994530941723212145818 + ln_k) / lambda);
}

void hauser_correction(const Scoring& sc, const int8_t* seq, int len, std::vector<int8_t>& out) {
	std::vector<float> f((size_t)len, 0.0f);
	int scores[20];
	std::memset(scores, 0, sizeof scores);
	auto add = [&](int l) { for (int i = 0; i < 20; ++i) scores[i] += sc.score(l, i); };
	auto sub = [&](int l) { for (int i = 0; i < 20; ++i) scores[i] -= sc.score(l, i); };
	auto at = [&](unsigned i) { return (int)(seq[i] & 31); };
	const unsigned window = 40, l = (unsigned)len, window_half = std::min(window / 2, l - 1);
	unsigned n = 0, h = 0, m = 0, t = 0;
	auto emit = [&](unsigned pos) {
		const int r = at(pos);
		if (r < 20) f[pos] = (float)sc.background_scores[r] - float(scores[r] - sc.score(r, r)) / (n - 1);
	};
	while (n < window_half && h < l) { ++n; add(at(h)); ++h; }
	while (n < (window + 1) && h < l) { ++n; add(at(h)); emit(m); ++h; ++m; }
	while (h < l) { add(at(h)); sub(at(t)); emit(m); ++h; ++t; ++m; }
	while (m < l && n > (window_half + 1)) { --n; sub(at(t)); emit(m); ++t; ++m; }
	while (m < l) { emit(m); ++m; }
	out.clear();
	out.reserve((size_t)len + 32);
	for (float x : f) out.push_back(int8_t(x < 0.0f ? x - 0.5f : x + 0.5f));
	out.insert(out.end(), 32, 0);
}

}  // namespace dmnd
