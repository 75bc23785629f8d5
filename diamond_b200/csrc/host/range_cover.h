// Step function over read positions behind --range-culling: shared by the legacy extension pipeline (legacy.inc, inside the library) and
// the join of a blocked run (cli.cpp), which cull per query range in the same way (output/target_culling.h:110-160).
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <iterator>
#include <map>

// IntervalPartition (util/geo/interval_partition.h): a step function over read positions -- how many inserted ranges cover a
// position, the smallest score among the first `cap` of them and the largest score -- behind --range-culling.  A node at position p
// holds the value of [p, next node).  insert() keeps the reference's construction order (value copied from the predecessor at the
// range start, the value in front of the last touched node restored at the range end).
class RangeCover {
public:
	struct Node { int64_t count = 0; int min_score = INT_MAX, max_score = 0; };
	explicit RangeCover(int64_t cap) : cap_(cap) { m_[0] = Node(); }
	void insert(int b, int e, int score) {
		auto i = m_.lower_bound(b);
		if (i == m_.end()) i = m_.emplace(b, Node()).first;
		else if (i->first != b) { auto prev = std::prev(i); i = m_.emplace(b, prev->second).first; }
		Node last;
		while (i != m_.end() && i->first < e) {
			last = i->second;
			Node& n = i->second;
			n = Node{ n.count + 1, n.count < cap_ ? std::min(n.min_score, score) : n.min_score, std::max(n.max_score, score) };
			++i;
		}
		if (i == m_.end() || i->first != e) m_[e] = last;
	}
	// letters of [b, e) that lie in segments accepted by `ok`
	template<typename F> int covered(int b, int e, F ok) const {
		auto i = m_.lower_bound(b), j = i;
		if (i == m_.end() || i->first != b) --i; else ++j;
		int c = 0;
		while (i != m_.end() && i->first < e) {
			const int sb = i->first, se = j == m_.end() ? INT_MAX : j->first;
			if (ok(i->second)) c += std::max(std::min(e, se) - std::max(b, sb), 0);
			i = j;
			if (j != m_.end()) ++j;
		}
		return c;
	}
	int covered_full(int b, int e) const { const int64_t cap = cap_; return covered(b, e, [cap](const Node& n) { return n.count >= cap; }); }
	int covered_max(int b, int e, int score) const { return covered(b, e, [score](const Node& n) { return n.max_score >= score; }); }
	int covered_min(int b, int e, int score) const { const int64_t cap = cap_; return covered(b, e, [cap, score](const Node& n) { return n.count >= cap && n.min_score >= score; }); }
private:
	std::map<int, Node> m_;
	int64_t cap_;
};
