// chaining.h -- ungapped x-drop segments and greedy chaining (see chaining.cpp for the reference map).
#pragma once
#include <climits>
#include <cstdint>
#include <vector>
#include "scoring.h"

namespace dmnd {

// DiagonalSegment (util/geo/diagonal_segment.h:22-140): ungapped segment starting at query i / subject j.
struct Segment {
	int i, j, len, score;
	int diag() const { return i - j; }
	int subject_last() const { return j + len - 1; }
	int query_last() const { return i + len - 1; }
	int subject_end() const { return j + len; }
	int query_end() const { return i + len; }
};

// ApproxHsp (util/hsp/approx_hsp.h:58-75) without the anchor.
struct Chain {
	int d_min = INT_MAX, d_max = INT_MIN, score = 0;
	int q_begin = 0, q_end = 0, s_begin = 0, s_end = 0;
};

// `query`/`subject` point at letter 0 of the sequences inside their blocks (delimiters on both sides).
Segment xdrop_ungapped(const Scoring& sc, const int8_t* query, const int8_t* cbs, const int8_t* subject, int qa, int sa);

// Chaining::run (chaining/greedy_align.cpp:482-497): `segs` sorted by (diag, j).
void chain_segments(const Scoring& sc, const int8_t* query, int qlen, const int8_t* subject, int slen,
                    const std::vector<Segment>& segs, std::vector<Chain>& out);

}  // namespace dmnd
