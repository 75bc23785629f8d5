// cli.cpp -- `dmnd-b200 blastp`: the reference's command-line surface for the hot path (run/main.cpp:73-209,
// basic/config.cpp:216-648) reduced to the flags the path understands; everything else is rejected, not ignored.
// FASTA in (data/fasta/fasta_file.cpp), block images as data/string_set.h:26-78, BLAST tabular out
// (output/blast_tab_format.cpp:234-293,652,694; util/text_buffer.h:224-246; util/string/string.h:87-92).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <zlib.h>
#include <iterator>
#include <unordered_map>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>
#include <string>
#include <algorithm>
#include "../../../include/dmnd_b200.h"
#include "range_cover.h"

namespace {

struct SeqBlock {
	std::vector<int8_t> letters;
	std::vector<int64_t> limits;
	std::vector<std::string> ids, titles;  // id = title up to the first blank (Util::Seq::id_delimiters)
	std::vector<uint32_t> oid;             // ordinal of every sequence in its file when the block was reordered (length_sort); empty = the block's own order
	std::vector<uint64_t> rec_begin;       // FASTA input: byte offset of every record's '>' in the (inflated) file, then the file size: block cuts (-b)
	SeqBlock() : letters(DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER) { limits.push_back(DMND_PERIMETER_PADDING); }
	void finish() { letters.insert(letters.end(), DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER); }
	uint32_t size() const { return (uint32_t)ids.size(); }
};

// Whole input file as text; gzip-compressed files are inflated, anything else is read as it is (gzread is transparent), like the
// reference's input layer (util/io/compressed_stream.cpp) for .gz queries and databases in FASTA format.
std::string slurp_text(const std::string& path) {
	gzFile g = path.empty() ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");  // no file name = standard input, as the reference reads its queries then
	if (!g) throw std::runtime_error("Error opening file " + path);
	std::string out;
	char buf[1 << 16];
	int n;
	while ((n = gzread(g, buf, sizeof buf)) > 0) out.append(buf, (size_t)n);
	const bool bad = n < 0;
	gzclose(g);
	if (bad) throw std::runtime_error("Error reading file " + path);
	return out;
}

// Block::length_sorted (data/block/block.cpp:229-255): longest sequence first, ties by DESCENDING block id (std::greater on (length, id)).
// The reference sorts the query and the reference block this way when min_length_ratio is set (run/double_indexed.cpp:112-115,727-731):
// queries are then reported in this order, and the block ids that break ranking ties are the sorted ones.
// `chunks` (optional): first sequence of every query block the reference would load (ascending, starting with 0): each is sorted on its own.
void length_sort(SeqBlock& b, const std::vector<uint32_t>* chunks = nullptr) {
	const uint32_t n = b.size();
	std::vector<std::pair<int64_t, uint32_t>> key(n);
	for (uint32_t i = 0; i < n; ++i) key[i] = { b.limits[i + 1] - b.limits[i] - 1, i };
	if (!chunks) std::sort(key.begin(), key.end(), std::greater<std::pair<int64_t, uint32_t>>());
	else for (size_t c = 0; c + 1 < chunks->size(); ++c) std::sort(key.begin() + (*chunks)[c], key.begin() + (*chunks)[c + 1], std::greater<std::pair<int64_t, uint32_t>>());
	SeqBlock s;
	s.letters.reserve(b.letters.size());
	for (uint32_t i = 0; i < n; ++i) {
		const uint32_t j = key[i].second;
		s.letters.insert(s.letters.end(), b.letters.begin() + b.limits[j], b.letters.begin() + b.limits[j + 1]);  // the letters and their delimiter
		s.limits.push_back((int64_t)s.letters.size());
		s.ids.push_back(std::move(b.ids[j]));
		s.oid.push_back(b.oid.empty() ? j : b.oid[j]);
		if (!b.titles.empty()) s.titles.push_back(std::move(b.titles[j]));
	}
	s.finish();
	b = std::move(s);
}

int8_t encode(char c) {  // basic/value.cpp:26-41 with amino_acid_traits (stats/stats.cpp:41): "UO-" -> mask
	static int8_t table[256];
	static bool init = false;
	if (!init) {
		std::memset(table, -1, sizeof table);
		const char* a = "ARNDCQEGHILKMFPSTWYVBJZX*_";
		for (int i = 0; a[i]; ++i) { table[(unsigned char)a[i]] = (int8_t)i; table[(unsigned char)tolower(a[i])] = (int8_t)i; }
		for (const char* m = "UO-"; *m; ++m) { table[(unsigned char)*m] = 23; table[(unsigned char)tolower(*m)] = 23; }
		init = true;
	}
	const int8_t v = table[(unsigned char)c];
	if (v < 0) throw std::runtime_error(std::string("Invalid character in sequence: '") + c + "'");
	return v;
}

void read_fasta(const std::string& path, SeqBlock& b) {
	const std::string text = slurp_text(path);
	std::istringstream f(text);
	std::string line;
	bool open = false;
	uint64_t line_begin = 0;
	auto close_seq = [&] {
		if (!open) return;
		b.letters.push_back((int8_t)DMND_DELIMITER);
		b.limits.push_back((int64_t)b.letters.size());
		open = false;
	};
	while (std::getline(f, line)) {
		const uint64_t this_line = line_begin;
		line_begin += line.size() + 1;
		if (!line.empty() && line.back() == '\r') line.pop_back();
		if (line.empty()) continue;
		if (line[0] == '>') {
			close_seq();
			b.rec_begin.push_back(this_line);
			size_t e = 1;
			while (e < line.size() && !strchr(" \t\x01", line[e])) ++e;  // Util::Seq::id_delimiters
			b.ids.push_back(line.substr(1, e - 1));
			b.titles.push_back(line.substr(1));
			open = true;
		}
		else {
			if (!open) throw std::runtime_error("FASTA format error: sequence data before the first header in " + path);
			for (char c : line) b.letters.push_back(encode(c));
		}
	}
	close_seq();
	b.rec_begin.push_back((uint64_t)text.size());
	b.finish();
}

// ---- blastx: DNA queries, translated into their six reading frames (util/sequence/translate.h:58-100, basic/basic.cpp:86-139,
// data/block/block.cpp:86-100).  The query block holds, for query s, the contexts 6s .. 6s+5: frames +1 +2 +3 of the read, then
// -1 -2 -3 of its reverse complement; every stretch between two stop codons that is shorter than min_orf_len is X-ed out.
struct DnaQueries {
	std::vector<std::string> ids, titles;
	std::vector<int32_t> len;  // nucleotides
	std::vector<std::string> qual; // FASTQ quality strings (empty for FASTA input): the qqual / full_qqual fields
	std::vector<std::string> dna;  // the reads as the reference prints them (nucleotide_traits alphabet: everything but ACGT is N)
};

int8_t encode_dna(char c) {  // nucleotide_traits (stats/stats.cpp:42): "ACGTN", everything in "MRWSYKVHDBX" reads as N
	switch (toupper((unsigned char)c)) {
	case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4;
	case 'M': case 'R': case 'W': case 'S': case 'Y': case 'K': case 'V': case 'H': case 'D': case 'B': case 'X': return 4;
	default: throw std::runtime_error(std::string("Invalid character in sequence: '") + c + "'");
	}
}

struct TranslateOpts { int strand_mask = 63, min_orf = 0, gencode = 1, frame_shift = 0; };  // --strand, --min-orf (0 = Config::min_orf_len's rule), --query-gencode, -F

// NCBI translation tables (the ids --query-gencode accepts, basic/basic.cpp:86-113), TCAG order; nullptr = no such table
const char* genetic_code(int id) {
	static const char* codes[27] = { nullptr,
		"FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",
		"FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		"FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG", "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		nullptr, nullptr,
		"FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		"FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		"FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG", "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
		nullptr,
		"FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		nullptr, nullptr, nullptr, nullptr,
		"FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG", "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
		"FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG",
		"FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG", "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG" };
	return id >= 1 && id <= 26 ? codes[id] : nullptr;
}

struct CodonTable {  // Translator::init(id): a codon with N is X unless its first two bases fix the amino acid
	int8_t fwd[5][5][5], rev[5][5][5];
	explicit CodonTable(int id) {
		const char* code = genetic_code(id);
		if (!code) throw std::runtime_error("Invalid genetic code id.");
		static const unsigned idx[4] = { 2, 1, 3, 0 }, comp[4] = { 3, 2, 1, 0 };
		for (unsigned i = 0; i < 5; ++i) for (unsigned j = 0; j < 5; ++j) for (unsigned k = 0; k < 5; ++k) {
			if (i == 4 || j == 4 || k == 4) { fwd[i][j][k] = rev[i][j][k] = 23; continue; }
			fwd[i][j][k] = encode(code[idx[i] * 16 + idx[j] * 4 + idx[k]]);
			rev[i][j][k] = encode(code[idx[comp[i]] * 16 + idx[comp[j]] * 4 + idx[comp[k]]]);
		}
		for (unsigned i = 0; i < 4; ++i) for (unsigned j = 0; j < 4; ++j) {
			if (std::count(fwd[i][j], fwd[i][j] + 4, fwd[i][j][0]) == 4) fwd[i][j][4] = fwd[i][j][0];
			if (std::count(rev[i][j], rev[i][j] + 4, rev[i][j][0]) == 4) rev[i][j][4] = rev[i][j][0];
		}
	}
};

void push_translated(const std::vector<int8_t>& dna, SeqBlock& b, const TranslateOpts& to) {
	static const CodonTable ct(to.gencode);  // one table per run
	const size_t L = dna.size();
	std::vector<int8_t> fr[6];
	if (L >= 3) {
		for (int f = 0; f < 3; ++f) {
			const size_t n = (L - (size_t)f) / 3;
			fr[f].resize(n); fr[f + 3].resize(n);
			for (size_t i = 0; i < n; ++i) {
				const size_t p = 3 * i + (size_t)f;            // codon start on the forward strand
				fr[f][i] = ct.fwd[dna[p]][dna[p + 1]][dna[p + 2]];
				const size_t r = L - 3 - p;                    // the reverse strand's codon i of frame f, read back to front
				fr[f + 3][i] = ct.rev[dna[r + 2]][dna[r + 1]][dna[r]];
			}
		}
	}
	const size_t l0 = fr[0].size();
	const size_t min_len = to.min_orf > 0 ? (size_t)to.min_orf : ((l0 < 30 || to.frame_shift != 0) ? 1 : (l0 < 100 ? 20 : 40));  // Config::min_orf_len (basic/config.h:413-424): no ORF masking in frameshift mode
	for (int f = 0; f < 6; ++f) {
		std::vector<int8_t>& v = fr[f];
		if (!((to.strand_mask >> f) & 1)) std::fill(v.begin(), v.end(), (int8_t)23);  // a strand that is not searched: the frame stays, all X (block.cpp:95-96)
		else for (size_t begin = 0;;) {  // Util::Seq::find_orfs (util/sequence/sequence.cpp:180-197)
			size_t it = begin;
			while (it < v.size() && v[it] != 24) ++it;
			if (it - begin < min_len) std::fill(v.begin() + (ptrdiff_t)begin, v.begin() + (ptrdiff_t)it, (int8_t)23);
			if (it >= v.size()) break;
			begin = it + 1;
		}
		b.letters.insert(b.letters.end(), v.begin(), v.end());
		b.letters.push_back((int8_t)DMND_DELIMITER);
		b.limits.push_back((int64_t)b.letters.size());
	}
}

void read_dna_fasta(const std::string& path, DnaQueries& dq, SeqBlock& b, const TranslateOpts& to) {
	std::istringstream f(slurp_text(path));
	std::string line;
	std::vector<int8_t> dna;
	bool open = false;
	auto close_seq = [&] {
		if (!open) return;
		dq.len.push_back((int32_t)dna.size());
		{ std::string t(dna.size(), 'N'); for (size_t k = 0; k < dna.size(); ++k) t[k] = "ACGTN"[dna[k]]; dq.dna.push_back(std::move(t)); }
		push_translated(dna, b, to);
		dna.clear();
		open = false;
	};
	bool fastq = false, first = true;
	while (std::getline(f, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		if (line.empty()) continue;
		if (first) { fastq = line[0] == '@'; first = false; }
		if (fastq) {  // FASTQ: @title, sequence line(s), +, quality line(s); the qualities are not used on this path
			if (line[0] != '@') throw std::runtime_error("FASTQ format error: missing @ in " + path);
			close_seq();
			size_t e = 1;
			while (e < line.size() && !strchr(" \t\x01", line[e])) ++e;
			dq.ids.push_back(line.substr(1, e - 1));
			dq.titles.push_back(line.substr(1) + "\n");  // the reference's FASTQ reader keeps a newline at the end of the title (data/fasta/parser.h:247): qtitle, -f 0 and -f 5 show it
			open = true;
			// FastqTokenizer::read_record (data/fasta/parser.h:238-270): sequence lines up to the '+' line, then quality lines until they
			// are as long as the sequence
			std::string l2;
			size_t len = 0, qlen = 0;
			for (;;) {
				if (!std::getline(f, l2)) throw std::runtime_error("Malformed FASTQ record in " + path);
				if (!l2.empty() && l2.back() == '\r') l2.pop_back();
				if (l2.empty()) throw std::runtime_error("Malformed FASTQ record in " + path);
				if (l2[0] == '+') break;
				len += l2.size();
				for (char c : l2) dna.push_back(encode_dna(c));
			}
			std::string qual;
			while (std::getline(f, l2)) {
				if (!l2.empty() && l2.back() == '\r') l2.pop_back();
				qlen += l2.size();
				qual += l2;
				if (qlen >= len) break;
			}
			dq.qual.resize(dq.ids.size());
			dq.qual.back() = std::move(qual);
			continue;
		}
		if (line[0] == '>') {
			close_seq();
			size_t e = 1;
			while (e < line.size() && !strchr(" \t\x01", line[e])) ++e;
			dq.ids.push_back(line.substr(1, e - 1));
			dq.titles.push_back(line.substr(1));
			open = true;
		}
		else {
			if (!open) throw std::runtime_error("FASTA format error: sequence data before the first header in " + path);
			for (char c : line) dna.push_back(encode_dna(c));
		}
	}
	close_seq();
	b.finish();
}

// ---- DIAMOND database files (legacy/dmnd/dmnd.h:28-66, dmnd.cpp:50-120,224-234,319-327): 40-byte header, a size-prefixed second
// header with a 128-bit hash, one record per sequence (0xff, letters with bit 7 = tantan soft mask, 0xff, title, 0), then the
// position array {u64 offset, u32 length, u32 0} with a terminating entry.  Little endian.
constexpr uint64_t DMND_MAGIC = 0x24af8a415ee186dull;
constexpr uint32_t DMND_BUILD = 182, DMND_DB_VERSION = 3;  // Const::build_version (basic/const.h:25), ReferenceHeader::current_db_version_prot

bool is_dmnd(const std::string& path) {
	std::ifstream f(path, std::ios::binary);
	uint64_t m = 0;
	return f.read((char*)&m, 8) && m == DMND_MAGIC;
}

void read_dmnd(const std::string& path, SeqBlock& b) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("Error opening file " + path);
	std::vector<char> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	auto u64 = [&](size_t o) { uint64_t v; std::memcpy(&v, d.data() + o, 8); return v; };
	auto u32 = [&](size_t o) { uint32_t v; std::memcpy(&v, d.data() + o, 4); return v; };
	if (d.size() < 96 || u64(0) != DMND_MAGIC) throw std::runtime_error("Database file is not a DIAMOND database.");
	if (u32(12) < 2 || u32(12) > 3) throw std::runtime_error("Database was built with an unsupported database format version.");
	const uint64_t nseq = u64(16), pos_off = u64(32), size = d.size();
	// header fields are untrusted: every bound is checked without arithmetic that could wrap
	if (pos_off < 96 || pos_off > size || (size - pos_off) / 16 < 1 || nseq > (size - pos_off) / 16 - 1) throw std::runtime_error("Database file is truncated.");
	size_t total = 0;
	for (uint64_t i = 0; i < nseq; ++i) total += (size_t)u32(pos_off + 16 * i + 8) + 1;
	if (total > size) throw std::runtime_error("Database file format error.");
	b.letters.reserve(b.letters.size() + total);
	for (uint64_t i = 0; i < nseq; ++i) {
		const uint64_t pos = u64(pos_off + 16 * i), len = u32(pos_off + 16 * i + 8);
		// record = 0xff, len letters, 0xff, title, NUL -- all of it before the position array
		if (pos < 96 || pos >= pos_off || len > pos_off - pos || pos_off - pos - len < 3 || (unsigned char)d[pos] != 0xff || (unsigned char)d[pos + 1 + len] != 0xff)
			throw std::runtime_error("Database file format error.");
		const size_t at = b.letters.size();
		b.letters.resize(at + len);
		const char* src = d.data() + pos + 1;
		for (uint64_t k = 0; k < len; ++k) b.letters[at + k] = (int8_t)(src[k] & 0x7f);  // the soft-mask bit is recomputed by the pipeline (same algorithm)
		b.letters.push_back((int8_t)DMND_DELIMITER);
		b.limits.push_back((int64_t)b.letters.size());
		const char* title = d.data() + pos + 2 + len;
		const size_t room = (size_t)(pos_off - (pos + 2 + len));
		const void* nul = std::memchr(title, 0, room);
		if (!nul) throw std::runtime_error("Database file format error.");
		const size_t tlen = (size_t)((const char*)nul - title);
		b.titles.emplace_back(title, tlen);
		size_t e = 0;
		while (e < tlen && !strchr(" \t\x01", title[e])) ++e;
		b.ids.emplace_back(title, e);
	}
	b.finish();
}

// MurmurHash3_x64_128 (Austin Appleby, public domain) with a 128-bit seed, as lib/murmurhash is called from make_db (dmnd.cpp:304-308)
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t fmix64(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
void murmur3_x64_128(const void* key, int len, uint64_t h[2]) {
	const uint8_t* data = (const uint8_t*)key;
	const int nblocks = len / 16;
	uint64_t h1 = h[0], h2 = h[1];
	const uint64_t c1 = 0x87c37b91114253d5ull, c2 = 0x4cf5ad432745937full;
	for (int i = 0; i < nblocks; ++i) {
		uint64_t k1, k2;
		std::memcpy(&k1, data + 16 * i, 8); std::memcpy(&k2, data + 16 * i + 8, 8);
		k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
		h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
		k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
		h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
	}
	const uint8_t* tail = data + nblocks * 16;
	uint64_t k1 = 0, k2 = 0;
	switch (len & 15) {
	case 15: k2 ^= (uint64_t)tail[14] << 48; [[fallthrough]];
	case 14: k2 ^= (uint64_t)tail[13] << 40; [[fallthrough]];
	case 13: k2 ^= (uint64_t)tail[12] << 32; [[fallthrough]];
	case 12: k2 ^= (uint64_t)tail[11] << 24; [[fallthrough]];
	case 11: k2 ^= (uint64_t)tail[10] << 16; [[fallthrough]];
	case 10: k2 ^= (uint64_t)tail[9] << 8; [[fallthrough]];
	case 9: k2 ^= (uint64_t)tail[8]; k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; [[fallthrough]];
	case 8: k1 ^= (uint64_t)tail[7] << 56; [[fallthrough]];
	case 7: k1 ^= (uint64_t)tail[6] << 48; [[fallthrough]];
	case 6: k1 ^= (uint64_t)tail[5] << 40; [[fallthrough]];
	case 5: k1 ^= (uint64_t)tail[4] << 32; [[fallthrough]];
	case 4: k1 ^= (uint64_t)tail[3] << 24; [[fallthrough]];
	case 3: k1 ^= (uint64_t)tail[2] << 16; [[fallthrough]];
	case 2: k1 ^= (uint64_t)tail[1] << 8; [[fallthrough]];
	case 1: k1 ^= (uint64_t)tail[0]; k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
	}
	h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
	h1 += h2; h2 += h1;
	h1 = fmix64(h1); h2 = fmix64(h2);
	h1 += h2; h2 += h1;
	h[0] = h1; h[1] = h2;
}

// `makedb` (DatabaseFile::make_db, dmnd.cpp:236-370 without taxonomy): FASTA in, .dmnd out; unless --masking 0 the letters carry the
// tantan soft-mask bit, computed by the library (dmnd_block_mask on the resident block).
int make_db(const std::string& in, std::string out_path, bool masking) {
	if (out_path.size() < 5 || out_path.substr(out_path.size() - 5) != ".dmnd") out_path += ".dmnd";
	SeqBlock b;
	read_fasta(in, b);
	for (uint32_t i = 0; i < b.size(); ++i) if (b.limits[i + 1] - b.limits[i] - 1 == 0) throw std::runtime_error("File format error: sequence of length 0");
	std::vector<uint8_t> soft(b.letters.size(), 0);
	if (masking && b.size()) {
		dmnd_search_opts o; dmnd_search_opts_default(&o);
		dmnd_params params;
		dmnd_ctx* ctx = nullptr;
		dmnd_block* blk = nullptr;
		uint64_t n = 0;
		if (dmnd_params_init(&o, &params) || dmnd_create(0, &params, &ctx) || dmnd_block_upload(ctx, b.letters.data(), b.letters.size(), b.limits.data(), b.size(), &blk)
		    || dmnd_block_mask(ctx, blk, DMND_MASK_TANTAN, 0, b.size(), &n))
			throw std::runtime_error(dmnd_last_error());
		std::vector<uint64_t> pos((size_t)n);
		if (n && dmnd_block_mask_fetch(ctx, pos.data(), pos.size())) throw std::runtime_error(dmnd_last_error());
		for (uint64_t p : pos) soft[(size_t)p] = 0x80;
		dmnd_block_free(ctx, blk); dmnd_destroy(ctx);
	}
	std::string rec;
	std::vector<std::pair<uint64_t, uint32_t>> pos_array;
	uint64_t offset = 96, letters = 0, hash[2] = { 0, 0 };
	for (uint32_t i = 0; i < b.size(); ++i) {
		const size_t beg = (size_t)b.limits[i], len = (size_t)(b.limits[i + 1] - b.limits[i] - 1);
		pos_array.emplace_back(offset, (uint32_t)len);
		const size_t r0 = rec.size();
		rec += '\xff';
		for (size_t k = 0; k < len; ++k) rec += (char)((uint8_t)b.letters[beg + k] | soft[beg + k]);
		rec += '\xff';
		rec += b.titles[i]; rec += '\0';
		murmur3_x64_128(rec.data() + r0 + 1, (int)len, hash);
		murmur3_x64_128(b.titles[i].data(), (int)b.titles[i].size(), hash);
		letters += len;
		offset += len + b.titles[i].size() + 3;
	}
	pos_array.emplace_back(offset, 0u);
	FILE* f = fopen(out_path.c_str(), "wb");
	if (!f) throw std::runtime_error("Error opening file " + out_path);
	const uint64_t nseq = b.size(), h2size = 48, zero = 0;
	fwrite(&DMND_MAGIC, 8, 1, f); fwrite(&DMND_BUILD, 4, 1, f); fwrite(&DMND_DB_VERSION, 4, 1, f); fwrite(&nseq, 8, 1, f); fwrite(&letters, 8, 1, f); fwrite(&offset, 8, 1, f);
	fwrite(&h2size, 8, 1, f); fwrite(hash, 16, 1, f);
	for (int k = 0; k < 4; ++k) fwrite(&zero, 8, 1, f);
	fwrite(rec.data(), 1, rec.size(), f);
	for (const auto& pr : pos_array) { const uint32_t z = 0; fwrite(&pr.first, 8, 1, f); fwrite(&pr.second, 4, 1, f); fwrite(&z, 4, 1, f); }
	fclose(f);
	fprintf(stderr, "Database sequences  %llu\nDatabase letters  %llu\nDatabase hash  ", (unsigned long long)nseq, (unsigned long long)letters);
	for (int k = 0; k < 16; ++k) fprintf(stderr, "%02x", ((const uint8_t*)hash)[k]);
	fprintf(stderr, "\n");
	return 0;
}


// ---- DIAMOND alignment archive (legacy/daa/daa_file.h:31-97): the second header as it lies in the file
struct DaaHeader2 {
	uint64_t diamond_build, db_seqs, db_seqs_used, db_letters, flags, query_records;
	int32_t mode, gap_open, gap_extend, reward, penalty, reserved1, reserved2, reserved3;
	double k, lambda, evalue, reserved5;
	char score_matrix[16];
	uint64_t block_size[256];
	char block_type[256];
};
constexpr uint64_t DAA_MAGIC = 0x3c0e53476d3ee36bull;

// `view`: a DAA file back into the structures the writers print from (legacy/daa/daa_record.cpp:30-107, view.cpp): queries (unpacked, DNA
// reads translated into their six frames), the target dictionary as a reference block without letters, one dmnd_match per stored
// alignment with the statistics re-derived from its transcript the way HspContext::parse does (basic/hssp.cpp:52-100: a frameshift
// mark counts as a column), e-value and bit score recomputed from the raw score and the database size in the header.
void load_daa(const std::string& path, bool* translated, SeqBlock& q, DnaQueries& dq, SeqBlock& r, std::vector<dmnd_match>& matches, std::vector<uint8_t>& transcripts,
              DaaHeader2& h2, const int8_t* score32) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("Error opening file " + path);
	std::string d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
	uint64_t h1[2];
	if (d.size() < sizeof h1 + sizeof h2) throw std::runtime_error("Input file is not a DAA file.");
	std::memcpy(h1, d.data(), sizeof h1);
	if (h1[0] != DAA_MAGIC) throw std::runtime_error("Input file is not a DAA file.");
	if (h1[1] > 1) throw std::runtime_error("DAA version requires later version of DIAMOND.");
	std::memcpy(&h2, d.data() + sizeof h1, sizeof h2);
	if (h2.block_size[0] == 0) throw std::runtime_error("Invalid DAA file. DIAMOND run has probably not completed successfully.");
	if (h2.mode != 2 && h2.mode != 3) throw std::runtime_error("view: only blastp and blastx archives are implemented");
	*translated = h2.mode == 3;
	const size_t base = sizeof h1 + sizeof h2;
	// the dictionary: names, then lengths
	size_t p = base + (size_t)h2.block_size[0];
	for (uint64_t i = 0; i < h2.db_seqs_used; ++i) {
		const size_t e = d.find('\0', p);
		if (e == std::string::npos) throw std::runtime_error("Invalid DAA file (target names).");
		r.ids.push_back(d.substr(p, e - p));
		r.titles.push_back(r.ids.back());
		p = e + 1;
	}
	if (p + 4 * h2.db_seqs_used > d.size()) throw std::runtime_error("Invalid DAA file (target lengths).");
	for (uint64_t i = 0; i < h2.db_seqs_used; ++i) {
		uint32_t l; std::memcpy(&l, d.data() + p + 4 * i, 4);
		r.letters.insert(r.letters.end(), (size_t)l, (int8_t)23);  // (the archive does not hold the targets: their letters come out of the transcripts)
		r.letters.push_back((int8_t)DMND_DELIMITER);
		r.limits.push_back((int64_t)r.letters.size());
	}
	r.finish();
	const TranslateOpts to{ 63, 0, 1, 1 };  // (the reference's view translates the stored read as it is: no ORF masking)
	auto rd = [&](size_t& at, unsigned code) -> uint32_t {  // read_packed: 1, 2 or 4 bytes
		uint32_t v = 0;
		const size_t nb = code == 0 ? 1 : code == 1 ? 2 : 4;
		if (at + nb > d.size()) throw std::runtime_error("Invalid DAA file (record).");
		std::memcpy(&v, d.data() + at, nb);
		at += nb;
		return v;
	};
	p = base;
	for (uint32_t qi = 0;; ++qi) {
		uint32_t size; std::memcpy(&size, d.data() + p, 4);
		if (size == 0) break;
		const size_t rec_end = p + 4 + size;
		if (rec_end > base + h2.block_size[0]) throw std::runtime_error("Invalid DAA file (query record).");
		size_t at = p + 4;
		uint32_t qlen; std::memcpy(&qlen, d.data() + at, 4); at += 4;
		const size_t ne = d.find('\0', at);
		if (ne == std::string::npos || ne >= rec_end) throw std::runtime_error("Invalid DAA file (query name).");
		const std::string name = d.substr(at, ne - at);
		at = ne + 1;
		const bool has_n = (d[at++] & 1) != 0;
		const unsigned bits = *translated ? (has_n ? 3u : 2u) : 5u;
		std::vector<int8_t> seq(qlen);
		{
			unsigned acc = 0, nb = 0; uint32_t l = 0;
			const size_t nbytes = ((size_t)qlen * bits + 7) / 8;
			if (at + nbytes > rec_end) throw std::runtime_error("Invalid DAA file (query sequence).");
			for (size_t b = 0; b < nbytes; ++b) {
				acc |= (unsigned)(uint8_t)d[at + b] << nb; nb += 8;
				while (nb >= bits && l < qlen) { seq[l++] = (int8_t)(acc & ((1u << bits) - 1)); nb -= bits; acc >>= bits; }
			}
			at += nbytes;
		}
		for (int8_t& c : seq) if (*translated ? c > 4 : c > 25) c = *translated ? 4 : 23;  // (a damaged file: keep the letters inside their alphabet)
		if (*translated) {
			dq.ids.push_back(name); dq.titles.push_back(name); dq.len.push_back((int32_t)qlen);
			std::string t(qlen, 'N'); for (uint32_t k = 0; k < qlen; ++k) t[k] = "ACGTN"[seq[k] > 4 ? 4 : seq[k]];
			dq.dna.push_back(std::move(t));
			push_translated(seq, q, to);
		}
		else {
			q.ids.push_back(name); q.titles.push_back(name);
			q.letters.insert(q.letters.end(), seq.begin(), seq.end());
			q.letters.push_back((int8_t)DMND_DELIMITER);
			q.limits.push_back((int64_t)q.letters.size());
		}
		while (at < rec_end) {
			dmnd_match x;
			std::memset(&x, 0, sizeof x);
			uint32_t dict; std::memcpy(&dict, d.data() + at, 4); at += 4;
			if (at + 5 > rec_end || dict >= r.size()) throw std::runtime_error("Invalid DAA file (match record).");
			const uint8_t flag = (uint8_t)d[at++];
			x.target = dict;
			x.score = (int32_t)rd(at, flag & 3u);
			const uint32_t qb = rd(at, (flag >> 2) & 3u);
			x.t_begin = (int32_t)rd(at, (flag >> 4) & 3u);
			int frame = 0, off = 0, pos = (int)qb;
			if (*translated) {  // IntermediateRecord::frame (output/output_format.cpp:79-84): the oriented begin on the read gives strand and frame
				const int in_strand = (flag & 64) ? (int)qlen - 1 - (int)qb : (int)qb;
				off = in_strand % 3; frame = ((flag & 64) ? 3 : 0) + off; pos = in_strand / 3;
			}
			x.query = *translated ? 6u * qi + (uint32_t)frame : qi;
			x.q_begin = pos;
			x.transcript_off = transcripts.size();
			const uint32_t c0 = *translated ? 6u * qi + (uint32_t)(frame / 3 * 3) : qi;
			int tpos = x.t_begin;
			unsigned gap_run = 0;
			bool shifted = false;
			for (;; ++at) {
				if (at >= rec_end) throw std::runtime_error("Invalid DAA file (transcript).");
				const uint8_t c = (uint8_t)d[at];
				if (c == 0) { ++at; break; }
				const unsigned op = c >> 6, val = c & 63u;
				if ((op == 2 && val > 25) || (op == 3 && val > 27)) throw std::runtime_error("Invalid DAA file (transcript letter).");
				auto qletter = [&]() -> int {
					const int64_t fl = q.limits[c0 + (uint32_t)off + 1] - q.limits[c0 + (uint32_t)off] - 1;
					if (pos < 0 || pos >= fl) throw std::runtime_error("Invalid DAA file (alignment outside its query).");
					return q.letters[(size_t)q.limits[c0 + (uint32_t)off] + (size_t)pos] & 31;
				};
				if (op == 0) for (unsigned k = 0; k < val; ++k) { qletter(); transcripts.push_back(0x00); ++x.identities; ++x.positives; ++x.length; ++pos; ++tpos; gap_run = 0; }
				else if (op == 1) for (unsigned k = 0; k < val; ++k) { qletter(); transcripts.push_back(0x40); if (gap_run++ == 0) ++x.gap_openings; ++x.gaps; ++x.length; ++pos; }
				else if (op == 2) { transcripts.push_back(c); if (gap_run++ == 0) ++x.gap_openings; ++x.gaps; ++x.length; ++tpos; }
				else if (val == 26 || val == 27) {  // frameshift: a column of its own for parse(), the frame of the following letters changes
					transcripts.push_back(val == 27 ? (uint8_t)DMND_TR_FRAMESHIFT_FWD : (uint8_t)DMND_TR_FRAMESHIFT_REV);
					++x.length; shifted = true;
					if (val == 27) { if (++off == 3) { off = 0; ++pos; } } else { if (--off < 0) { off = 2; --pos; } }
				}
				else { transcripts.push_back(c); ++x.mismatches; ++x.length; if (score32[qletter() * 32 + (int)val] > 0) ++x.positives; ++pos; ++tpos; gap_run = 0; }
			}
			x.transcript_len = (uint32_t)(transcripts.size() - x.transcript_off);
			x.q_end = pos; x.t_end = tpos;
			{
				const int64_t fl = q.limits[c0 + (uint32_t)off + 1] - q.limits[c0 + (uint32_t)off] - 1, bl = q.limits[x.query + 1] - q.limits[x.query] - 1;
				if (x.q_begin < 0 || x.q_begin > bl || pos < 0 || pos > fl + 1 || tpos > r.limits[dict + 1] - r.limits[dict] - 1) throw std::runtime_error("Invalid DAA file (alignment outside its sequences).");
			}
			if (*translated && shifted) x.reserved = 1u + (uint32_t)(frame / 3 * 3 + off);
			const uint32_t tlen = (uint32_t)(r.limits[dict + 1] - r.limits[dict] - 1);
			const uint32_t c_first = *translated ? 6u * qi : qi;  // (the reference's view takes the length of the FIRST frame, daa_record.cpp:84 -- a search takes the aligned frame's)
			const uint32_t ql = (uint32_t)(q.limits[c_first + 1] - q.limits[c_first] - 1);
			dmnd_alignment_stats(x.score, ql, tlen, h2.db_letters, &x.evalue, &x.bit_score);
			matches.push_back(x);
		}
		p = rec_end;
	}
	q.finish();
}

int format_double(double x, char* p, size_t n) {  // util/string/string.h:87-92
	if (x >= 100.0) return snprintf(p, n, "%lli", (long long)std::floor(x));
	const long long i = std::llround(x * 10.0);
	return snprintf(p, n, "%lli.%lli", i / 10, i % 10);
}

[[noreturn]] void usage(const char* msg) {
	fprintf(stderr, "Error: %s\nusage: dmnd-b200 makedb --in DB.faa -d DB\n"
	                "       dmnd-b200 blastp|blastx -d DB[.dmnd|.faa] [-q QUERIES] [-o OUT] [--fast|--mid-sensitive|--sensitive|--more-sensitive|--very-sensitive|--ultra-sensitive]\n"
	                "                [-p N] [-c N] [-b G] [-k N | --top P] [-e X | --min-score B] [--id P | --approx-id P] [--query-cover P] [--subject-cover P] [--no-self-hits]\n"
	                "                [--comp-based-stats 0|1] [--masking 0|1] [--motif-masking 0|1] [-F 15 [--range-culling] | --long-reads] [--strand both|plus|minus] [--min-orf N] [--query-gencode N]\n"
	                "                [-f 6|104 [fields] [--unal 0|1] [--header simple] | -f 0 | -f 5 | -f 100 | -f sam | -f paf] [--salltitles|--sallseqid] [--ext banded-fast|banded-slow] [--compress 0|1] [--log]\n"
	                "       dmnd-b200 view -a FILE.daa [-o OUT] [-f ...]\n       dmnd-b200 dbinfo -d DB\n", msg);
	exit(1);
}

}  // namespace

int main(int argc, char** argv) {
	try {
		if (argc < 2) usage("missing command");
		const std::string cmd = argv[1];
		if (cmd == "version") { printf("dmnd-b200 (%s) for diamond 2.2.2 blastp hot path\n", dmnd_backend()); return 0; }
		if (cmd == "makedb") {  // diamond makedb --in X.faa -d OUT [--masking 0]
			std::string in, db;
			const bool masking = true;
			for (int i = 2; i < argc; ++i) {
				const std::string a = argv[i];
				auto val = [&]() -> const char* { if (i + 1 >= argc) usage(("missing value for " + a).c_str()); return argv[++i]; };
				if (a == "--in") in = val();
				else if (a == "-d" || a == "--db") db = val();
				else if (a == "--masking") usage("Option is not permitted for this workflow: masking");  // as the reference answers
				else if (a == "-p" || a == "--threads") val();
				else if (a == "--quiet") {}
				else usage(("unsupported option " + a).c_str());
			}
			if (in.empty() || db.empty()) usage("makedb needs --in and -d");
			return make_db(in, db, masking);
		}
		if (cmd == "dbinfo") {  // diamond dbinfo -d DB: the header of a DIAMOND database (legacy/dmnd/dmnd.h:28-66)
			std::string db;
			for (int i = 2; i < argc; ++i) { const std::string a = argv[i]; if ((a == "-d" || a == "--db") && i + 1 < argc) db = argv[++i]; else if (a != "--quiet") usage(("unsupported option " + a).c_str()); }
			if (db.empty()) usage("dbinfo needs -d");
			if (!is_dmnd(db) && is_dmnd(db + ".dmnd")) db += ".dmnd";
			std::ifstream f(db, std::ios::binary);
			unsigned char h[40];
			uint64_t magic = 0, seqs = 0, letters = 0;
			uint32_t build = 0, version = 0;
			if (!f.read((char*)h, sizeof h)) throw std::runtime_error("Database file is not a DIAMOND database.");
			std::memcpy(&magic, h, 8); std::memcpy(&build, h + 8, 4); std::memcpy(&version, h + 12, 4); std::memcpy(&seqs, h + 16, 8); std::memcpy(&letters, h + 24, 8);
			if (magic != DMND_MAGIC) throw std::runtime_error("Database file is not a DIAMOND database.");
			printf("%23s  %s\n%23s  %u\n%23s  %u\n%23s  %llu\n%23s  %llu\n", "Database type", "Diamond database", "Database format version", version, "Diamond build", build,
			       "Sequences", (unsigned long long)seqs, "Letters", (unsigned long long)letters);
			return 0;
		}
		if (cmd != "blastp" && cmd != "blastx" && cmd != "view") usage("only blastp, blastx, view and makedb are implemented");
		const bool view_mode = cmd == "view";  // diamond view -a FILE.daa -o OUT [-f ...]: the stored alignments through the same writers
		std::string daa_in;
		bool translated = cmd == "blastx";
		if (view_mode) {  // the archive's header says whether its queries are reads (the options below depend on it)
			for (int i = 2; i + 1 < argc; ++i) if (!strcmp(argv[i], "-a") || !strcmp(argv[i], "--daa")) daa_in = argv[i + 1];
			if (daa_in.empty()) usage("view needs -a FILE.daa");
			if (!std::ifstream(daa_in) && std::ifstream(daa_in + ".daa")) daa_in += ".daa";  // auto_append_extension_if_exists
			std::ifstream hf(daa_in, std::ios::binary);
			uint64_t h1[2] = { 0, 0 };
			DaaHeader2 hh;
			if (!hf.read((char*)h1, sizeof h1) || !hf.read((char*)&hh, sizeof hh) || h1[0] != DAA_MAGIC) usage("Input file is not a DAA file.");
			translated = hh.mode == 3;
		}
		dmnd_search_opts o;
		dmnd_search_opts_default(&o);
		if (translated) o.query_contexts = 6;
		o.sensitivity = 1;  // like the reference: no sensitivity flag = Sensitivity::DEFAULT, --fast = Sensitivity::FAST
		std::string qf, df, of;
		bool log = false, motif_set = false;
		std::vector<std::string> fields;
		bool pairwise = false, paf = false, sam = false, xml = false, daa = false, json = false, k_set = false, top_set = false, unal = false;
		int strand_mask = 63, min_orf = 0, gencode = 1;
		bool header_simple = false, long_reads = false, no_self_hits = false, gz_out = false, no_auto_append = false, salltitles = false, sallseqid = false, xml_blord = false, no_parse_seqids = false, sam_qlen = false;
		uint64_t daa_build = 182;  // Const::build_version of the reference release this path follows
		std::string matrix_name = "blosum62";  // config.matrix as given (the XML header quotes it)
		double block_size = 0.0;  // -b: reference block size in 10^9 letters (0 = the mode's default: 2.0, 0.4 from --very-sensitive on; run/double_indexed.cpp:792-795)
		for (int i = 2; i < argc; ++i) {
			std::string a = argv[i];
			// a short option with its value attached (-p4, -c1, -k0, -f0, -e10000), as the reference's parser accepts it
			const char* attached = nullptr;
			if (a.size() > 2 && a[0] == '-' && a[1] != '-' && strchr("pckefqdobFat", a[1])) { attached = argv[i] + 2; a = a.substr(0, 2); }
			auto val = [&]() -> const char* { if (attached) return attached; if (i + 1 >= argc) usage(("missing value for " + a).c_str()); return argv[++i]; };
			if (view_mode && (a == "-a" || a == "--daa")) val();
			else if (a == "-q" || a == "--query") qf = val();
			else if (a == "-d" || a == "--db") df = val();
			else if (a == "-o" || a == "--out") of = val();
			else if (a == "--fast") o.sensitivity = 0;
			else if (a == "--mid-sensitive") o.sensitivity = 2;
			else if (a == "--sensitive") o.sensitivity = 3;
			else if (a == "--more-sensitive") o.sensitivity = 4;
			else if (a == "--very-sensitive") o.sensitivity = 5;
			else if (a == "--ultra-sensitive") o.sensitivity = 6;
			else if (a == "-p" || a == "--threads") o.threads = atoi(val());
			else if (a == "-c" || a == "--index-chunks") o.index_chunks = atoi(val());
			else if (a == "-b" || a == "--block-size") { block_size = atof(val()); if (!(block_size > 0.0)) usage("--block-size must be positive"); }
			else if (a == "-k" || a == "--max-target-seqs") { o.max_target_seqs = atoi(val()); k_set = true; }
			else if (a == "-e" || a == "--evalue") o.max_evalue = atof(val());
			else if (a == "--top") { o.top_percent = atof(val()); top_set = true; if (o.top_percent < 0.0 || o.top_percent > 100.0) usage("Allowed value range for --top is between 0.0 and 100.0"); }
			else if (a == "--comp-based-stats") o.comp_based_stats = atoi(val());
			else if (a == "--masking") {  // masking/masking.cpp:42-48: 0/none, 1/tantan (seg is not part of this build)
				const std::string v = val();
				if (v == "0" || v == "none") o.masking = 0; else if (v == "1" || v == "tantan") o.masking = 1; else usage("--masking must be 0, none, 1 or tantan (seg is not implemented)");
			}
			else if (a == "--motif-masking") {  // search/setup.cpp:322-336
				const std::string v = val();
				if (v == "0") o.motif_masking = 0; else if (v == "1") o.motif_masking = 1; else usage("Permitted values for --motif-masking: 0, 1");
				motif_set = true;
			}
			else if (a == "-f" || a == "--outfmt") {  // -f 6 [field ...]  (output/blast_tab_format.cpp:41-118; the 12 default fields + the transcript fields)
				const std::string fmt = val();
				if (fmt == "0") { pairwise = true; continue; }
				if (fmt == "paf" || fmt == "103") { paf = true; continue; }
				if (fmt == "sam" || fmt == "101") { sam = true; continue; }
				if (fmt == "xml" || fmt == "5") { xml = true; continue; }
				if (fmt == "daa" || fmt == "100") { daa = true; continue; }
				if (fmt == "json-flat" || fmt == "104") json = true;  // the tabular fields as a JSON array of flat objects (TabularFormat(true), blast_tab_format.cpp:740-773,823-834)
				if (fmt != "6" && fmt != "tab" && !json) usage("only -f 6 [fields], -f 0, -f 5 (xml), -f 100 (daa), -f sam and -f paf are implemented");
				while (i + 1 < argc && argv[i + 1][0] != '-') {
					const std::string f = argv[++i];
					static const char* known[] = { "qseqid", "sseqid", "pident", "length", "mismatch", "gapopen", "qstart", "qend", "sstart", "send", "evalue", "bitscore",
					                               "cigar", "btop", "qseq_gapped", "sseq_gapped", "score", "gaps", "nident", "qlen", "slen",
					                               "qtitle", "stitle", "positive", "ppos", "qcovhsp", "scovhsp", "qframe", "qstrand", "qseq", "sseq",
					                               "sallseqid", "salltitles", "full_sseq", "full_qseq", "qnum", "snum", "hspnum", "qseq_translated", "normalized_nident", "qqual", "full_qqual", "approx_pident" };
					bool ok = false;
					for (const char* k : known) ok |= f == k;
					if (!ok) usage(("unsupported output field " + f).c_str());
					fields.push_back(f);
				}
			}
			else if (a == "--query-parallel-limit") val();  // a threading knob of the reference's extension stage: no effect on the result (its own goldens agree)
			else if (a == "--algo") {  // 0 / double-indexed is what this path implements; 1 / query-indexed is the reference's other join order over the same
				const std::string v = val();  // seeds, defined to give the same alignments (src/test: diamond-test-blastp-query-indexed.out == ...-more-sensitive.out)
				if (v != "0" && v != "1" && v != "double-indexed" && v != "query-indexed") usage("--algo must be 0, 1, double-indexed or query-indexed");
			}
			else if (a == "--header") {  // TabularFormat::header_format (output/blast_tab_format.cpp:627-641): 0 = none, simple = the field keys; the
				std::string v = "verbose";  // verbose form (no value) quotes the invocation and the long field descriptions and is not built
				if (i + 1 < argc && argv[i + 1][0] != '-') v = argv[++i];
				if (v == "0") header_simple = false; else if (v == "simple") header_simple = true; else usage("--header: only 0 and simple are implemented");
			}
			else if (a == "--unal") { const std::string v = val(); if (v != "0" && v != "1") usage("--unal must be 0 or 1"); unal = v == "1"; }
			else if (a == "--strand") {  // frame_mask(), data/sequence_file.cpp:286-294
				const std::string v = val();
				if (v == "both") strand_mask = 63; else if (v == "plus") strand_mask = 7; else if (v == "minus") strand_mask = 56; else usage("Invalid value for parameter --strand");
			}
			else if (a == "--min-orf" || a == "-l") { min_orf = atoi(val()); if (min_orf < 0) usage("--min-orf must not be negative"); }
			else if (a == "--query-gencode") gencode = atoi(val());
			else if (a == "--range-culling") o.range_culling = 1;
			else if (a == "--range-cover") o.range_cover = atof(val());  // config.query_range_cover (default 50)
			else if (a == "--long-reads") long_reads = true;  // basic/config.cpp:679-686
			else if (a == "-F" || a == "--frameshift") { o.frame_shift = atoi(val()); if (o.frame_shift <= 0) usage("--frameshift needs a positive penalty (the reference's usual value is 15)"); }
			else if (a == "--log") log = true;
			else if (a == "--quiet" || a == "--verbose" || a == "-v" || a == "--ignore-warnings") {}
			else if (a == "--tmpdir" || a == "-t" || a == "--parallel-tmpdir") val();  // (no temporary files on this path: hits and per-block results stay in memory)
			else if (a == "--daa-build-version") daa_build = (uint64_t)atoll(val());  // config.daa_build_version: the build number a DAA header quotes (daa_file.h:45)
			else if (a == "--no-auto-append") no_auto_append = true;  // basic/config.cpp:765
			// options that wrappers (e.g. the Galaxy tool) always spell out: accepted at their default value only, anything else is refused
			else if (a == "--compress") { const std::string v = val(); if (v == "1") gz_out = true; else if (v != "0") usage("--compress: 0 (none) and 1 (gzip) are implemented"); }
			else if (a == "--matrix") { std::string v = val(); matrix_name = v; for (char& c : v) c = (char)toupper((unsigned char)c); if (v != "BLOSUM62") usage("--matrix: only BLOSUM62 (gap penalties 11/1) is implemented"); }
			else if (a == "--id") o.min_id = atof(val());  // basic/config.cpp:263,300-301: report filters, applied inside the extension (align/culling.cpp:144-184)
			else if (a == "--xml-blord-format") xml_blord = true;  // XML: <Hit_id>gnl|BL_ORD_ID|ordinal</Hit_id>, the whole title line in Hit_def (xml_format.cpp:44-49)
			else if (a == "--no-parse-seqids") no_parse_seqids = true;  // XML: Hit_accession = the id as it is (xml_format.cpp:60)
			else if (a == "--sam-query-len") sam_qlen = true;  // SAM: a ZQ:i:<query length> field (sam_format.cpp:136-137)
			else if (a == "--salltitles") salltitles = true;  // SAM reference names and DAA dictionary entries carry every full title (basic/config.cpp; sam_format.cpp:100, daa_record.cpp:27)
			else if (a == "--sallseqid") sallseqid = true;   // DAA dictionary entries carry every sequence id of a merged record
			else if (a == "--no-self-hits") no_self_hits = true;  // basic/config.cpp:312
			else if (a == "--min-score") o.min_bit_score = atof(val());  // basic/config.cpp:299: overrides the e-value setting
			else if (a == "--query-cover") o.query_cover = atof(val());
			else if (a == "--subject-cover") o.subject_cover = atof(val());
			else if (a == "--approx-id") o.approx_min_id = atof(val());
			else if (a == "--ext") {  // search/setup.cpp:377-384
				const std::string v = val();
				if (v == "banded-fast") o.ext_mode = 1; else if (v == "banded-slow") o.ext_mode = 2; else usage("--ext: banded-fast and banded-slow are implemented (full, global and none are not on this path)");
			}
			else if (a == "--gapopen") { if (atoi(val()) != 11) usage("--gapopen: only 11 (BLOSUM62's default, with --gapextend 1) is implemented"); }
			else if (a == "--gapextend") { if (atoi(val()) != 1) usage("--gapextend: only 1 (with --gapopen 11) is implemented"); }
			else if (a == "--max-hsps") { if (std::string(val()) != "1") usage("--max-hsps: only 1 is implemented"); }
			else usage(("unsupported option " + a).c_str());
		}
		if (view_mode) {
			if (daa) usage("view: the output format is one of 0, 5, 6, sam, paf");
			if (o.min_id != 0.0 || o.query_cover != 0.0 || o.subject_cover != 0.0 || o.approx_min_id != 0.0 || o.min_bit_score != 0.0 || unal || block_size != 0.0 || o.frame_shift || no_self_hits)
				usage("view prints the stored alignments: search options do not apply");
		}
		else if (df.empty()) usage("-d is required");  // (no -q: the queries come from standard input; no -o: the output goes to standard output)
		if (of.empty() && (daa || gz_out)) usage("-f 100 and --compress need an output file (-o)");
		if (daa && gz_out) usage("Compression is not supported for DAA format.");  // basic/config.cpp:726-727
		if (daa && !no_auto_append && (of.size() < 4 || of.compare(of.size() - 4, 4, ".daa") != 0)) of += ".daa";  // auto_append_extension, basic/config.cpp:725-730
		if (k_set && top_set) usage("--top and --max-target-seqs are mutually exclusive.");  // basic/config.cpp:674-675
		if (long_reads) {  // --long-reads = --range-culling --top 10 -F 15 (each only where not given)
			o.range_culling = 1;
			if (!top_set) { o.top_percent = 10.0; top_set = true; }
			if (o.frame_shift == 0) o.frame_shift = 15;
		}
		if (o.range_culling && o.frame_shift == 0) usage("Query range culling is only supported in frameshift alignment mode (option -F).");  // basic/config.cpp:824-825
		if (pairwise || paf || sam || xml || daa) o.want_transcript = 1;  // both formats ask for HspValues::TRANSCRIPT (output/output_format.h:205-216)
		for (const std::string& f : fields) if (f == "cigar" || f == "btop" || f == "qseq_gapped" || f == "sseq_gapped" || f == "positive" || f == "ppos" || f == "sseq" || f == "qseq_translated") o.want_transcript = 1;  // HspValues::TRANSCRIPT
		for (const std::string& f : fields) if (f == "qseq_translated" && !translated) usage("Output field only supported for translated search.");  // output/blast_tab_format.cpp:685-686
		if (!motif_set) o.motif_masking = dmnd_mode_motif_masking(o.sensitivity);  // the mode's default (traits.motif_masking, search/setup.cpp:322-325)
		if (o.comp_based_stats != 0 && o.comp_based_stats != 1) usage("--comp-based-stats must be 0 or 1");
		SeqBlock q, r;
		DnaQueries dq;
		if (!translated && (strand_mask != 63 || min_orf != 0 || gencode != 1)) usage("--strand, --min-orf and --query-gencode belong to blastx");
		if (no_self_hits && translated) usage("--no-self-hits option is not supported in blastx mode.");  // basic/config.cpp:677-678
		if (o.frame_shift && !translated) usage("Frameshift alignments are only supported for translated searches.");  // basic/config.cpp:822-823
		const bool fshift = o.frame_shift != 0;
		if (fshift) {
			o.want_transcript = 1;  // output/output_format.cpp:256-257
		}
		dmnd_params params;
		if (dmnd_params_init(&o, &params)) throw std::runtime_error(dmnd_last_error());
		std::vector<dmnd_match> view_matches;
		std::vector<uint8_t> view_transcripts;
		DaaHeader2 view_header;
		if (view_mode) load_daa(daa_in, &translated, q, dq, r, view_matches, view_transcripts, view_header, params.score);
		else {
			if (translated) read_dna_fasta(qf, dq, q, TranslateOpts{ strand_mask, min_orf, gencode, o.frame_shift });
			else read_fasta(qf, q);
			if (is_dmnd(df) || (!std::ifstream(df) && is_dmnd(df + ".dmnd"))) read_dmnd(is_dmnd(df) ? df : df + ".dmnd", r);  // -d takes a DIAMOND database or a FASTA file, like the reference
			else read_fasta(df, r);
		}
		const uint32_t nq_block = translated ? (uint32_t)dq.ids.size() * 6u : q.size();
		dmnd_ctx* ctx = nullptr;
		if (!view_mode && dmnd_create(0, &params, &ctx)) throw std::runtime_error(dmnd_last_error());
		// ---- reference blocks (run/double_indexed.cpp:218-244): a database beyond the block size is searched block by block with the
		// e-values of the WHOLE database, and the per-block results are joined per query (output/join_blocks.cpp).  Block cuts follow the
		// reference's loaders: a .dmnd database by letters (SequenceFile::load_twopass, data/sequence_file.cpp:214-240: sequences are
		// added while the block holds fewer than block-size letters), a FASTA database by bytes of the file (load_parallel :104-189 with
		// FastaFile::raw_chunk: records up to the first record boundary at or after block-size bytes).
		if (block_size == 0.0) block_size = o.sensitivity >= 5 ? 0.4 : 2.0;
		const uint64_t max_letters = (uint64_t)(block_size * 1e9);
		auto block_cuts = [&](const SeqBlock& b) {
			std::vector<uint32_t> c{ 0 };
			const uint32_t nb = b.size();
			uint32_t f = 0;
			while (f < nb) {
				uint32_t e = f;
				if (!b.rec_begin.empty()) { while (e < nb && b.rec_begin[e] - b.rec_begin[f] < max_letters) ++e; }
				else { uint64_t letters = 0; while (e < nb && letters < max_letters) { letters += (uint64_t)(b.limits[e + 1] - b.limits[e] - 1); ++e; } }
				c.push_back(e);
				f = e;
			}
			if (c.size() == 1) c.push_back(0);
			return c;
		};
		const std::vector<uint32_t> cuts = block_cuts(r);
		const size_t nblocks = cuts.size() - 1;
		bool mutual = false;
		if (!translated && o.query_cover >= 50.0 && o.query_cover == o.subject_cover) {  // min_length_ratio is set (run/config.cpp:156-159): both blocks are searched length-sorted
			mutual = true;
			// (the query file is loaded in blocks of the same size, each sorted on its own: only a block size below the query file's shows it)
			const std::vector<uint32_t> qcuts = block_cuts(q);
			length_sort(q, qcuts.size() > 2 ? &qcuts : nullptr);
			if (nblocks == 1) length_sort(r);  // (more blocks: every block is sorted on its own when it is searched, below)
		}
		std::vector<std::vector<uint32_t>> bperm(nblocks);   // mutual coverage with several blocks: sorted block id -> id within the unsorted block
		std::vector<std::vector<int64_t>> blimits(nblocks);  // ... and the limits of the sorted block image
		uint64_t db_letters = 0;
		for (uint32_t i = 0; i < r.size(); ++i) db_letters += (uint64_t)(r.limits[i + 1] - r.limits[i] - 1);
		// --no-self-hits: an alignment of a query with the target of the same title and the same letters is not reported (filter_hsp,
		// align/culling.cpp:166-168).  Titles live here, so the pairs are found here: self_of[target] = the query it is a copy of
		std::vector<uint32_t> self_of, self_targets;
		if (no_self_hits) {
			std::unordered_map<std::string, std::vector<uint32_t>> by_title;
			for (uint32_t i = 0; i < q.size(); ++i) by_title[q.titles[i]].push_back(i);
			self_of.assign(r.size(), UINT32_MAX);
			for (uint32_t t2 = 0; t2 < r.size(); ++t2) {
				auto it = by_title.find(r.titles[t2]);
				if (it == by_title.end()) continue;
				const int64_t tl = r.limits[t2 + 1] - r.limits[t2];
				for (uint32_t qi : it->second)
					if (q.limits[qi + 1] - q.limits[qi] == tl && std::equal(r.letters.begin() + r.limits[t2], r.letters.begin() + r.limits[t2 + 1], q.letters.begin() + q.limits[qi],
					                                                       [](int8_t a2, int8_t b2) { return (a2 & 31) == (b2 & 31); })) { self_of[t2] = qi; break; }
			}
		}
		auto self_for_block = [&](size_t bk) -> const uint32_t* {  // the block's own target numbers (its sort order, if it was length-sorted)
			if (!no_self_hits) return nullptr;
			self_targets.assign(q.size(), UINT32_MAX);
			for (uint32_t tl = 0; tl < cuts[bk + 1] - cuts[bk]; ++tl) {
				const uint32_t g = cuts[bk] + (bperm[bk].empty() ? tl : bperm[bk][tl]);
				if (self_of[g] != UINT32_MAX) self_targets[self_of[g]] = tl;
			}
			return self_targets.data();
		};
		std::vector<dmnd_result*> results(nblocks, nullptr);
		for (size_t bk = 0; bk < nblocks && !view_mode; ++bk) {
			if (nblocks == 1) {
				o.self_targets = self_for_block(0);
				if (dmnd_blastp(ctx, q.letters.data(), q.letters.size(), q.limits.data(), nq_block, r.letters.data(), r.letters.size(), r.limits.data(), r.size(), &o, &results[0]))
					throw std::runtime_error(dmnd_last_error());
				break;
			}
			const uint32_t f = cuts[bk], e = cuts[bk + 1];
			const int64_t a = r.limits[f], z = r.limits[e];
			std::vector<int8_t> bl((size_t)DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER);
			bl.insert(bl.end(), r.letters.begin() + a, r.letters.begin() + z);
			bl.insert(bl.end(), (size_t)DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER);
			std::vector<int64_t> lim((size_t)(e - f) + 1);
			for (uint32_t k = f; k <= e; ++k) lim[k - f] = r.limits[k] - a + DMND_PERIMETER_PADDING;
			if (mutual) {  // Block::length_sorted of this reference block (run/double_indexed.cpp:112-115)
				std::vector<std::pair<int64_t, uint32_t>> key((size_t)(e - f));
				for (uint32_t k = 0; k < e - f; ++k) key[k] = { lim[k + 1] - lim[k] - 1, k };
				std::sort(key.begin(), key.end(), std::greater<std::pair<int64_t, uint32_t>>());
				std::vector<int8_t> sl((size_t)DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER);
				std::vector<int64_t> slim{ (int64_t)DMND_PERIMETER_PADDING };
				for (const auto& kv : key) {
					sl.insert(sl.end(), bl.begin() + lim[kv.second], bl.begin() + lim[kv.second + 1]);
					slim.push_back((int64_t)sl.size());
					bperm[bk].push_back(kv.second);
				}
				sl.insert(sl.end(), (size_t)DMND_PERIMETER_PADDING, (int8_t)DMND_DELIMITER);
				bl.swap(sl); lim.swap(slim);
				blimits[bk] = lim;
			}
			dmnd_search_opts ob = o;
			ob.db_letters = db_letters;
			ob.self_targets = self_for_block(bk);
			if (dmnd_blastp(ctx, q.letters.data(), q.letters.size(), q.limits.data(), nq_block, bl.data(), bl.size(), lim.data(), e - f, &ob, &results[bk]))
				throw std::runtime_error(dmnd_last_error());
		}
		dmnd_result* res = results[0];
		size_t n = view_matches.size();
		const dmnd_match* m = view_mode ? view_matches.data() : dmnd_result_matches(res, &n);
		size_t ntr = 0;
		const uint8_t* tr = view_mode ? view_transcripts.data() : dmnd_result_transcripts(res, &ntr);
		if (fields.empty()) fields = { "qseqid", "sseqid", "pident", "length", "mismatch", "gapopen", "qstart", "qend", "sstart", "send", "evalue", "bitscore" };
		// (full_sseq prints the target as loaded -- the reference keeps an unmasked copy of the block for it -- while a protein query's full_qseq shows
		// its masked letters like every other sequence field: a copy of the reference letters is taken before the masked ones are patched in)
		std::vector<int8_t> r_unmasked;
		for (const std::string& f : fields) if (f == "full_sseq") r_unmasked = r.letters;
		// sequence-bearing fields print the MASKED letters, as the reference does (its blocks are masked in place)
		for (size_t bk = 0; bk < nblocks && !view_mode; ++bk)
			for (int side = 0; side < 2; ++side) {
				size_t nm = 0;
				const uint64_t* mp = dmnd_result_masked_positions(results[bk], side, &nm);
				std::vector<int8_t>& l = side ? r.letters : q.letters;
				const int64_t shift = (side && nblocks > 1) ? r.limits[cuts[bk]] - DMND_PERIMETER_PADDING : 0;  // block image position -> position in the whole database image
				for (size_t k = 0; k < nm; ++k) {
					int64_t pos = (int64_t)mp[k];
					if (side && !bperm[bk].empty()) {  // position in the sorted block image -> the same letter in the unsorted one
						const std::vector<int64_t>& sl = blimits[bk];
						const size_t i = (size_t)(std::upper_bound(sl.begin(), sl.end(), pos) - sl.begin()) - 1;
						pos = pos - sl[i] + (r.limits[cuts[bk] + bperm[bk][i]] - r.limits[cuts[bk]] + DMND_PERIMETER_PADDING);
					}
					l[(size_t)(pos + shift)] = 23;
				}
			}
		// join_blocks (output/join_blocks.cpp:120-265): per query a merge of the blocks' match lists, best e-value first (then score, then
		// database order -- JoinRecord::cmp_evalue; by score with --top), cut by GlobalCulling (output/target_culling.h:39-90): the first
		// target always, then max_target_seqs targets, or the targets within --top percent of the best bit score
		std::vector<dmnd_match> joined;
		std::vector<uint8_t> joined_tr;
		std::vector<uint32_t> joined_unal;
		if (nblocks > 1) {
			struct Cur { const dmnd_match* m; size_t n, i; const uint8_t* tr; uint32_t first; };
			std::vector<Cur> cur(nblocks);
			for (size_t bk = 0; bk < nblocks; ++bk) { cur[bk].m = dmnd_result_matches(results[bk], &cur[bk].n); cur[bk].i = 0; size_t x = 0; cur[bk].tr = dmnd_result_transcripts(results[bk], &x); cur[bk].first = cuts[bk]; }
			const uint32_t cx = translated ? 6u : 1u;
			const bool top_on = top_set && o.top_percent != 100.0;  // --top 100 unsets toppercent: every target, ranked by e-value (output/output_format.cpp:236-239)
			const bool by_score = top_on;
			auto oid = [&](size_t bk, uint32_t t) { return cuts[bk] + (bperm[bk].empty() ? t : bperm[bk][t]); };  // database order of a block's target
			auto before = [&](const dmnd_match& a, uint32_t oa, const dmnd_match& b, uint32_t ob) {  // a comes out of the heap before b
				if (!by_score && a.evalue != b.evalue) return a.evalue < b.evalue;
				if (a.score != b.score) return a.score > b.score;
				return oa < ob;
			};
			for (;;) {
				uint32_t src = UINT32_MAX;
				for (const Cur& c : cur) if (c.i < c.n) src = std::min(src, c.m[c.i].query / cx);
				if (src == UINT32_MAX) break;
				std::vector<size_t> end(nblocks);
				for (size_t bk = 0; bk < nblocks; ++bk) { size_t e = cur[bk].i; while (e < cur[bk].n && cur[bk].m[e].query / cx == src) ++e; end[bk] = e; }
				int64_t n_targets = 0;
				double top_bits = 0.0;
				// --range-culling: the join culls per query range as the extension did (RangeCulling::cull / add over the records' absolute_query_range,
				// output/target_culling.h:132-160): a record whose read range is >= 50 % covered by records already reported is skipped, never final
				RangeCover part(top_on ? 0 : ((o.max_target_seqs == 0 || top_set) ? INT64_MAX : (int64_t)o.max_target_seqs));
				for (;;) {
					size_t best = nblocks;
					for (size_t bk = 0; bk < nblocks; ++bk) {
						if (cur[bk].i >= end[bk]) continue;
						if (best == nblocks || before(cur[bk].m[cur[bk].i], oid(bk, cur[bk].m[cur[bk].i].target), cur[best].m[cur[best].i], oid(best, cur[best].m[cur[best].i].target))) best = bk;
					}
					if (best == nblocks) break;
					const dmnd_match& x = cur[best].m[cur[best].i];
					if (o.range_culling) {
						const int fr = (int)(x.query % 6), off = fr % 3, ef = x.reserved ? ((int)x.reserved - 1) % 3 : off;
						const int L = dq.len[x.query / 6], b_in = 3 * x.q_begin + off, e_in = 3 * x.q_end + ef;
						const int rb = fr < 3 ? b_in : L - e_in, re = fr < 3 ? e_in : L - b_in;
						const int c = top_on ? part.covered_max(rb, re, int((double)x.score / (1.0 - o.top_percent / 100.0))) : part.covered_full(rb, re);
						++cur[best].i;
						if (!((double)c / (double)(re - rb) * 100.0 < (o.range_cover > 0.0 ? o.range_cover : 50.0))) continue;  // config.query_range_cover
						part.insert(rb, re, x.score);
						dmnd_match y = x;
						y.target = oid(best, x.target);
						y.transcript_off = joined_tr.size();
						if (x.transcript_len) joined_tr.insert(joined_tr.end(), cur[best].tr + x.transcript_off, cur[best].tr + x.transcript_off + x.transcript_len);
						joined.push_back(y);
						continue;
					}
					if (n_targets > 0) {  // GlobalCulling::cull
						if (top_on) { if ((1.0 - x.bit_score / top_bits) * 100.0 > o.top_percent) break; }
						else if (n_targets >= (int64_t)((o.max_target_seqs == 0 || top_set) ? INT32_MAX : o.max_target_seqs)) break;
					}
					else top_bits = x.bit_score;
					dmnd_match y = x;
					y.target = oid(best, x.target);
					y.transcript_off = joined_tr.size();
					if (x.transcript_len) joined_tr.insert(joined_tr.end(), cur[best].tr + x.transcript_off, cur[best].tr + x.transcript_off + x.transcript_len);
					joined.push_back(y);
					++n_targets;
					++cur[best].i;
				}
				for (size_t bk = 0; bk < nblocks; ++bk) cur[bk].i = end[bk];
			}
			// a blocked run reports EVERY query without a joined record as unaligned (output/join_blocks.cpp:302-308,365-372: the join walks
			// the query ids between two records and after the last one) -- not only those with seed hits, as a one-block run does
			{
				const uint32_t nsrc = translated ? (uint32_t)dq.ids.size() : q.size();
				size_t ji = 0;
				for (uint32_t sq = 0; sq < nsrc; ++sq) {
					while (ji < joined.size() && joined[ji].query / cx < sq) ++ji;
					if (!(ji < joined.size() && joined[ji].query / cx == sq)) joined_unal.push_back(sq * cx);
				}
			}
			if (fshift) {
				// Blocked runs of the legacy pipeline travel through IntermediateRecords, and the join re-derives the alignment statistics from
				// the transcript (HspContext::parse, basic/hssp.cpp:52-100) instead of keeping the traceback's counters: `length` then counts
				// the frameshift marks as columns, and a gap run is not interrupted by a mark
				for (dmnd_match& y : joined) {
					const uint8_t* t = joined_tr.data() + y.transcript_off;
					y.length = y.identities = y.mismatches = y.gap_openings = y.gaps = 0;
					unsigned d = 0;
					for (uint32_t k = 0; k < y.transcript_len; ++k) {
						++y.length;
						if (t[k] == DMND_TR_FRAMESHIFT_FWD || t[k] == DMND_TR_FRAMESHIFT_REV) continue;
						const int op = t[k] >> 6;
						if (op == DMND_OP_MATCH) { ++y.identities; d = 0; }
						else if (op == DMND_OP_SUBSTITUTION) { ++y.mismatches; d = 0; }
						else { if (d == 0) ++y.gap_openings; ++d; ++y.gaps; }
					}
				}
			}
			m = joined.data(); n = joined.size(); tr = joined_tr.data();
		}
		// frameshift mode: the legacy pipeline visits EVERY query, also those without a seed hit (align/align.cpp:120-130,168-172), so every
		// query without an alignment is reported by the formats that report unaligned queries
		// -- but only inside a query BIN of the seed-hit buffer that holds at least one hit: align_queries loops over the bins
		// (search/hit_buffer.cpp:149-188, one bin per load), a bin without hits starts no worker, and the unaligned records of its
		// queries never appear.  Bins = SequenceSet::partition(query_bins, true, true) (data/sequence_set.cpp:57-75): query_bins =
		// max(round(threads / 8), 16; 64 for --ultra-sensitive) runs of whole queries of >= ceil(letters / query_bins) letters.
		std::vector<uint32_t> fs_unal;
		if (fshift) {
			const uint32_t nsrc = (uint32_t)dq.ids.size();
			std::vector<uint8_t> has_hits(nsrc, 0);
			for (size_t i = 0; i < n; ++i) has_hits[m[i].query / 6] = 1;
			{ size_t nu = 0; const uint32_t* u = dmnd_result_unaligned(res, &nu); for (size_t i = 0; i < nu; ++i) has_hits[u[i] / 6] = 1; }
			const unsigned n_part = std::max((unsigned)std::lround((double)o.threads / 8.0), o.sensitivity >= 6 ? 64u : 16u);
			int64_t letters = 0;
			for (uint32_t c = 0; c < 6 * nsrc; ++c) letters += q.limits[c + 1] - q.limits[c] - 1;
			const int64_t per_bin = (letters + n_part - 1) / n_part;
			std::vector<uint8_t> visited(nsrc, 0);
			for (uint32_t sq = 0; sq < nsrc;) {
				int64_t acc = 0;
				const uint32_t b0 = sq;
				while (sq < nsrc && acc < per_bin) { for (uint32_t c = 6 * sq; c < 6 * sq + 6; ++c) acc += q.limits[c + 1] - q.limits[c] - 1; ++sq; }
				bool any = false;
				for (uint32_t x = b0; x < sq; ++x) any |= has_hits[x] != 0;
				if (any) std::fill(visited.begin() + b0, visited.begin() + sq, (uint8_t)1);
			}
			size_t mi = 0;
			for (uint32_t sq = 0; sq < nsrc; ++sq) {
				while (mi < n && m[mi].query / 6 < sq) ++mi;
				if (!(mi < n && m[mi].query / 6 == sq) && visited[sq]) fs_unal.push_back(6 * sq);
			}
		}
		auto result_unaligned = [&](size_t* nu) -> const uint32_t* {
			if (view_mode) { *nu = 0; return nullptr; }  // (an archive holds aligned queries only)
			if (nblocks > 1) { *nu = joined_unal.size(); return joined_unal.data(); }  // (a blocked run: the join's rule, frameshift mode or not)
			if (fshift) { *nu = fs_unal.size(); return fs_unal.data(); }
			return dmnd_result_unaligned(res, nu);
		};
		// every title of a merged record ("\x01" or " >" between them, util/sequence/sequence.cpp:38), whole or cut to its id, joined by `sep` (OutputFormat::print_title)
		auto all_titles = [](const std::string& tt, bool ids_only, const char* sep) {
			std::string out_;
			for (size_t a = 0, k = 0; a <= tt.size(); ++k) {
				size_t e = a;
				while (e < tt.size() && tt[e] != '\x01' && !(tt[e] == ' ' && e + 1 < tt.size() && tt[e + 1] == '>')) ++e;
				if (k) out_ += sep;
				size_t ie = e;
				if (ids_only) { ie = a; while (ie < e && !strchr(" \a\b\f\n\r\t\v", tt[ie])) ++ie; }
				out_.append(tt, a, ie - a);
				if (e >= tt.size()) break;
				a = tt[e] == '\x01' ? e + 1 : e + 2;
			}
			return out_;
		};
		static const char* alphabet = "ARNDCQEGHILKMFPSTWYVBJZX*_";
		// Hsp::Iterator (basic/match.h:105-161) over a transcript: qat[k] = the query letter transcript byte k consumes (match, substitution,
		// insertion), -1 / -2 for a forward / reverse frameshift byte (the query position moves one nucleotide on / back: the frame of the
		// letters after it changes, TranslatedPosition::shift_forward / shift_back), 0 for a deletion.  qcur = position after byte k:
		// codon index * 3 + frame offset inside the strand (TranslatedPosition::in_strand).  Without frameshift bytes this is qs[qi++].
		std::vector<int> qat, qcur;
		auto walk_query = [&](const dmnd_match& x, const uint8_t* t) {
			qat.assign(x.transcript_len, 0); qcur.assign((size_t)x.transcript_len + 1, 0);
			const uint32_t c0 = x.query - (translated ? x.query % 3 : 0);  // first frame of the strand
			int off = translated ? (int)(x.query % 3) : 0, pos = x.q_begin;
			qcur[0] = 3 * pos + off;
			for (uint32_t k = 0; k < x.transcript_len; ++k) {
				const uint8_t b = t[k];
				if (b == DMND_TR_FRAMESHIFT_FWD) { qat[k] = -1; if (++off == 3) { off = 0; ++pos; } }
				else if (b == DMND_TR_FRAMESHIFT_REV) { qat[k] = -2; if (--off < 0) { off = 2; --pos; } }
				else if ((b >> 6) != DMND_OP_DELETION) { qat[k] = q.letters[(size_t)q.limits[c0 + (uint32_t)off] + (size_t)pos] & 31; ++pos; }
				qcur[k + 1] = 3 * pos + off;
			}
		};
		auto end_frame = [&](const dmnd_match& x) -> int { return x.reserved ? (int)x.reserved - 1 : (int)(x.query % 6); };  // frame the alignment ends in
		FILE* out = of.empty() ? stdout : fopen(of.c_str(), "wb");
		if (!out) throw std::runtime_error("Error opening file " + of);
		char buf[32];
		std::string line;
		if (sam) {
			// SamFormat (output/sam_format.cpp:30-148): header (the @PG line carries THIS program's command line), one record per match,
			// the "4 *" record for queries that had seed hits but no alignment
			line = "@HD\tVN:1.5\tSO:query\n@PG\tPN:DIAMOND\tVN:2.2.2\tCL:";
			for (int ai = 0; ai < argc; ++ai) { if (ai) line += ' '; line += argv[ai]; }
			const char* ms = translated ? "BlastX" : "BlastP";
			line += std::string("\n@mm\t") + ms + "\n@CO\t" + ms + "-like alignments\n@CO\tReporting AS: bitScore, ZR: rawScore, ZE: expected, ZI: percent identity, ZL: reference length, ZF: frame, ZS: query start DNA coordinate\n";
			fwrite(line.data(), 1, line.size(), out);
			const uint32_t cx = translated ? 6u : 1u;
			size_t nu = 0, u = 0;
			const uint32_t* unal_ids = result_unaligned(&nu);
			auto unaligned_to = [&](uint32_t src_end) {
				for (; u < nu && unal_ids[u] / cx < src_end; ++u) {
					line = (translated ? dq.ids[unal_ids[u] / cx] : q.ids[unal_ids[u]]) + "\t4\t*\t0\t255\t*\t*\t0\t0\t*\t*\n";
					fwrite(line.data(), 1, line.size(), out);
				}
			};
			for (size_t i = 0; i < n; ++i) {
				const dmnd_match& x = m[i];
				const uint32_t sq = x.query / cx;
				unaligned_to(sq);
				const uint8_t* t = tr + x.transcript_off;
				const int8_t* qs = q.letters.data() + q.limits[x.query];
				line = (translated ? dq.ids[sq] : q.ids[sq]) + "\t0\t" + (salltitles ? all_titles(r.titles[x.target], false, "<>") : r.ids[x.target]) + "\t" + std::to_string(x.t_begin + 1) + "\t255\t";
				{	// print_cigar: match and substitution are M
					uint32_t run = 0; int op = -1;
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						const int o2 = t[k] >> 6;
						const int c = t[k] == DMND_TR_FRAMESHIFT_FWD ? 3 : t[k] == DMND_TR_FRAMESHIFT_REV ? 4 : (o2 == DMND_OP_INSERTION) ? 1 : (o2 == DMND_OP_DELETION) ? 2 : 0;  // frameshifts: \ and /
						if (c == op) ++run; else { if (run) { line += std::to_string(run); line += "MID\\/"[op]; } run = 1; op = c; }
					}
					if (run) { line += std::to_string(run); line += "MID\\/"[op]; }
				}
				line += "\t*\t0\t0\t";
				// SEQ = query_range().length() letters of the frame the alignment BEGINS in (sam_format.cpp:114).  After a forward frameshift the range can
				// end one codon past that frame: the reference then prints the block's delimiter through alphabet[31], one of the bytes BEHIND its
				// 26-letter alphabet literal -- 'Y' in the binary gcc builds from its sources here (the "MRWSYKVHDBX" literal follows); mirrored as such
				for (int p2 = x.q_begin; p2 < x.q_end; ++p2) line += (qs[p2] & 31) == DMND_DELIMITER ? 'Y' : alphabet[qs[p2] & 31];
				const int fr = translated ? (int)(x.query % 6) : 0, off = fr % 3;
				const int64_t b_in = 3 * (int64_t)x.q_begin + off;
				const int64_t zs = !translated ? x.q_begin + 1 : (fr < 3 ? b_in + 1 : (int64_t)dq.len[sq] - b_in);
				line += "\t*\tAS:i:" + std::to_string((uint32_t)x.bit_score) + "\tNM:i:" + std::to_string(x.length - x.identities) + "\tZL:i:" + std::to_string(r.limits[x.target + 1] - r.limits[x.target] - 1)
					+ "\tZR:i:" + std::to_string(x.score) + "\tZE:f:";
				if (x.evalue == 0.0) line += "0.0"; else { snprintf(buf, sizeof buf, "%.2e", x.evalue); line += buf; }
				line += "\tZI:i:" + std::to_string(x.identities * 100 / x.length) + "\tZF:i:" + std::to_string(fr < 3 ? fr + 1 : -(off + 1)) + "\tZS:i:" + std::to_string(zs) + "\tMD:Z:";
				{	// print_md (sam_format.cpp:30-64)
					unsigned matches = 0, del = 0;
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						const int o2 = t[k] >> 6;
						if (o2 == DMND_OP_MATCH) { del = 0; ++matches; }
						else if (o2 == DMND_OP_SUBSTITUTION) {
							if (matches > 0) { line += std::to_string(matches); matches = 0; }
							else if (del > 0) { line += '0'; del = 0; }
							line += alphabet[t[k] & 63];
						}
						else if (o2 == DMND_OP_DELETION) {
							if (matches > 0) { line += std::to_string(matches); matches = 0; }
							if (del == 0) line += '^';
							line += alphabet[t[k] & 63];
							++del;
						}
					}
					if (matches > 0) line += std::to_string(matches);
				}
				if (sam_qlen) line += "\tZQ:i:" + std::to_string(translated ? (int64_t)dq.len[sq] : q.limits[sq + 1] - q.limits[sq] - 1);  // r.query.source().length()
				line += '\n';
				fwrite(line.data(), 1, line.size(), out);
			}
			unaligned_to(UINT32_MAX);
			n = 0;
		}
		if (paf) {
			// PAFFormat (output/paf_format.cpp:24-68): one line per match, 0-based closed coordinates; queries without a match are
			// reported too when the extension stage saw them, i.e. when they had seed hits (Output::Flags::DEFAULT_REPORT_UNALIGNED, align/align.cpp:167-181), in query order
			const uint32_t nsrc = translated ? (uint32_t)dq.ids.size() : q.size();
			size_t i = 0, nu = 0, u = 0;
			const uint32_t* unal = result_unaligned(&nu);  // queries with seed hits and no alignment; queries without seed hits print nothing
			for (uint32_t s = 0; s < nsrc; ++s) {
				const std::string& qid = translated ? dq.ids[s] : q.ids[s];
				const size_t i0 = i;
				while (i < n && (translated ? m[i].query / 6 : m[i].query) == s) ++i;
				line.clear();
				while (u < nu && (translated ? unal[u] / 6 : unal[u]) < s) ++u;
				if (i == i0 && u < nu && (translated ? unal[u] / 6 : unal[u]) == s) line = qid + "\t4\t*\t0\t255\t*\t*\t0\t0\t*\t*\n";
				for (size_t k = i0; k < i; ++k) {
					const dmnd_match& x = m[k];
					int64_t qlen = q.limits[x.query + 1] - q.limits[x.query] - 1, qb = x.q_begin, qe = x.q_end;
					char strand = '+';
					if (translated) {  // query_source_range: TranslatedPosition::absolute_interval (basic/translated_position.h:121-127)
						const int fr = (int)(x.query % 6), off = fr % 3;
						qlen = dq.len[s];
						const int64_t b_in = 3 * (int64_t)x.q_begin + off, e_in = 3 * (int64_t)x.q_end + end_frame(x) % 3;  // (after a frameshift the end lies in its own frame)
						if (fr < 3) { qb = b_in; qe = e_in; } else { qb = qlen - e_in; qe = qlen - b_in; strand = '-'; }
					}
					line += qid + "\t" + std::to_string(qlen) + "\t" + std::to_string(qb) + "\t" + std::to_string(qe - 1) + "\t" + strand + "\t" + r.ids[x.target] + "\t"
						+ std::to_string(r.limits[x.target + 1] - r.limits[x.target] - 1) + "\t" + std::to_string(x.t_begin) + "\t" + std::to_string(x.t_end - 1) + "\t"
						+ std::to_string(x.identities) + "\t" + std::to_string(x.length) + "\t255\tAS:i:" + std::to_string((uint32_t)x.bit_score) + "\tZR:i:" + std::to_string(x.score) + "\tZE:f:";
					if (x.evalue == 0.0) line += "0.0"; else { snprintf(buf, sizeof buf, "%.2e", x.evalue); line += buf; }
					line += '\n';
				}
				fwrite(line.data(), 1, line.size(), out);
			}
			n = 0;
		}
		if (pairwise) {
			// PairwiseFormat (output/blast_pairwise_format.cpp:24-101): header once, an intro per aligned query, one record per match in
			// 60-column blocks; numbers go through TextBuffer::print(i, width), which keeps only `width` characters (util/text_buffer.h:248-254)
			const dmnd_params* pp = &params;
			auto put_num = [&](unsigned v, unsigned width) { char nb[24]; snprintf(nb, sizeof nb, "%*u", (int)width, v); line.append(nb, width); };
			line = "BLASTP 2.3.0+\n\n\n";  // the same header for blastx (print_header, blast_pairwise_format.cpp:97-101)
			fwrite(line.data(), 1, line.size(), out);
			size_t nu = 0, u = 0;
			const uint32_t* unal_ids = result_unaligned(&nu);
			const uint32_t cx = translated ? 6u : 1u;
			auto intro = [&](uint32_t sq) {
				return "Query= " + (translated ? dq.titles[sq] : q.titles[sq]) + "\n\nLength=" + std::to_string(translated ? (int64_t)dq.len[sq] : q.limits[sq + 1] - q.limits[sq] - 1) + "\n\n";
			};
			auto no_hits_upto = [&](uint32_t src_end) {  // DEFAULT_REPORT_UNALIGNED: queries with seed hits and no alignment (print_query_intro :88-95)
				for (; u < nu && unal_ids[u] / cx < src_end; ++u) {
					line = intro(unal_ids[u] / cx) + "\n***** No hits found *****\n\n\n";
					fwrite(line.data(), 1, line.size(), out);
				}
			};
			for (size_t i = 0; i < n; ++i) {
				const dmnd_match& x = m[i];
				const uint32_t sq = x.query / cx;
				no_hits_upto(sq);
				const uint8_t* t = tr + x.transcript_off;
				const int8_t* qs = q.letters.data() + q.limits[x.query];
				line.clear();
				if (i == 0 || m[i - 1].query / cx != sq) line += intro(sq);
				line += ">";
				{	// OutputFormat::print_title(out, title, true, true, " "): the titles of a merged record ("\x01" or " >" between them,
					// util/sequence/sequence.cpp:38) joined by one blank
					const std::string& tt = r.titles[x.target];
					for (size_t a = 0; a < tt.size();) {
						if (tt[a] == '\x01') { line += ' '; ++a; }
						else if (tt[a] == ' ' && a + 1 < tt.size() && tt[a + 1] == '>') { line += ' '; a += 2; }
						else line += tt[a++];
					}
				}
				line += "\nLength=" + std::to_string(r.limits[x.target + 1] - r.limits[x.target] - 1) + "\n\n";
				format_double(x.bit_score, buf, sizeof buf);
				line += " Score = "; line += buf; line += " bits (" + std::to_string(x.score) + "),  Expect = ";
				if (x.evalue == 0.0) line += "0.0"; else { snprintf(buf, sizeof buf, "%.2e", x.evalue); line += buf; }
				const unsigned len = (unsigned)x.length;
				line += "\n Identities = " + std::to_string(x.identities) + "/" + std::to_string(len) + " (" + std::to_string((unsigned)x.identities * 100u / len) + "%), Positives = "
					+ std::to_string(x.positives) + "/" + std::to_string(len) + " (" + std::to_string((unsigned)x.positives * 100u / len) + "%), Gaps = " + std::to_string(x.gaps) + "/"
					+ std::to_string(len) + " (" + std::to_string((unsigned)x.gaps * 100u / len) + "%)\n";
				// translated queries: the frame, and query positions on the read (TranslatedPosition::absolute / oriented_position,
				// blast_pairwise_format.cpp:41-58): a letter at frame position p sits at in-strand 3p + offset, mirrored on the reverse strand
				const int fr = translated ? (int)(x.query % 6) : 0, off = fr % 3;
				const int64_t L = translated ? dq.len[sq] : 0;
				if (translated) line += " Frame = " + std::to_string(fr < 3 ? fr + 1 : 2 - fr) + "\n";
				line += "\n";
				auto qpos_first = [&](int p) -> int64_t { return !translated ? p + 1 : (fr < 3 ? 3 * (int64_t)p + off + 1 : L - 3 * (int64_t)p - off); };  // the line's first letter
				auto qpos_end = [&](int p) -> int64_t { return !translated ? p : (fr < 3 ? 3 * (int64_t)p + off : L - 3 * (int64_t)p - off + 1); };       // after its last one (p = next position)
				const int64_t q_src_end = !translated ? x.q_end : (fr < 3 ? 3 * (int64_t)x.q_end + end_frame(x) % 3 : L - (3 * (int64_t)x.q_begin + off));
				const unsigned digits = (unsigned)std::max(std::ceil(std::log10((double)x.t_end)), std::ceil(std::log10((double)q_src_end)));
				(void)qpos_first; (void)qpos_end; (void)qs;
				walk_query(x, t);
				// line numbers from the in-strand position (3 * codon + frame offset): first letter TranslatedPosition::absolute + 1, after the
				// last one oriented_position(in_strand - 1) + 1 (blast_pairwise_format.cpp:54-63)
				auto in_first = [&](int in) -> int64_t { return !translated ? in / 3 + 1 : (fr < 3 ? (int64_t)in + 1 : L - in); };
				auto in_end = [&](int in) -> int64_t { return !translated ? in / 3 : (fr < 3 ? (int64_t)in : L - in + 1); };
				int si = x.t_begin;
				for (uint32_t k0 = 0; k0 < x.transcript_len; k0 += 60) {
					const uint32_t k1 = std::min<uint32_t>(k0 + 60, x.transcript_len);
					std::string ql, ml, sl;
					const int s0 = si;
					for (uint32_t k = k0; k < k1; ++k) {
						const int op = t[k] >> 6, sc = t[k] & 63;
						if (qat[k] < 0) { ql += qat[k] == -1 ? '\\' : '/'; sl += '-'; ml += ' '; }
						else if (op == DMND_OP_MATCH) { const char c = alphabet[qat[k]]; ql += c; ml += c; sl += c; ++si; }
						else if (op == DMND_OP_SUBSTITUTION) { const int a = qat[k]; ql += alphabet[a]; sl += alphabet[sc]; ml += pp->score[a * 32 + sc] > 0 ? '+' : ' '; ++si; }
						else if (op == DMND_OP_INSERTION) { ql += alphabet[qat[k]]; sl += '-'; ml += ' '; }
						else { ql += '-'; sl += alphabet[sc]; ml += ' '; ++si; }
					}
					line += "Query  "; put_num((unsigned)in_first(qcur[k0]), digits); line += "  " + ql + " " + std::to_string(in_end(qcur[k1])) + "\n";
					line.append(digits + 9, ' '); line += ml + "\n";
					line += "Sbjct  "; put_num((unsigned)s0 + 1, digits); line += "  " + sl + " " + std::to_string(si) + "\n\n";
				}
				fwrite(line.data(), 1, line.size(), out);
			}
			no_hits_upto(UINT32_MAX);
			n = 0;  // nothing left for the tabular writer
		}
		if (daa) {
			// DIAMOND alignment archive (legacy/daa/daa_file.h:31-97, daa_write.cpp:29-112, output/output.h:33-55, basic/packed_sequence.h,
			// basic/packed_transcript.h): two headers, one record per aligned query (length, id, the packed query -- 5 bits per residue, 2 or 3 per
			// nucleotide -- and its matches: dictionary id of the target, flag byte, score / oriented query begin / subject begin in 1, 2 or 4
			// bytes, the packed transcript), a zero, then the dictionary: target ids and lengths in order of first use.  The reference
			// numbers the dictionary in the order its threads reach the targets; this writer's order is that of a one-thread run (-p 1).
			if (nblocks > 1) throw std::runtime_error("-f 100 (DAA) is implemented for a database of one block");
			struct Header2 {
				uint64_t diamond_build, db_seqs, db_seqs_used, db_letters, flags, query_records;
				int32_t mode, gap_open, gap_extend, reward, penalty, reserved1, reserved2, reserved3;
				double k, lambda, evalue, reserved5;
				char score_matrix[16];
				uint64_t block_size[256];
				char block_type[256];
			} h2;
			std::memset(&h2, 0, sizeof h2);
			const uint64_t h1[2] = { 0x3c0e53476d3ee36bull, 1 };
			std::string body;
			auto put = [&](const void* p2, size_t nb) { body.append((const char*)p2, nb); };
			auto put_packed = [&](uint32_t v) { if (v <= 0xFFu) { const uint8_t b = (uint8_t)v; put(&b, 1); } else if (v <= 0xFFFFu) { const uint16_t b = (uint16_t)v; put(&b, 2); } else put(&v, 4); };
			auto len_flag = [](uint32_t v) -> unsigned { return v <= 0xFFu ? 0u : v <= 0xFFFFu ? 1u : 2u; };
			std::vector<uint32_t> dict_of(r.size(), UINT32_MAX), dict;
			const uint32_t cx = translated ? 6u : 1u;
			uint64_t n_queries = 0;
			size_t rec_pos = 0;
			for (size_t i = 0; i < n; ++i) {
				const dmnd_match& x = m[i];
				const uint32_t sq = x.query / cx;
				if (i == 0 || m[i - 1].query / cx != sq) {
					rec_pos = body.size();
					const uint32_t zero = 0, qlen = (uint32_t)(translated ? (int64_t)dq.len[sq] : q.limits[sq + 1] - q.limits[sq] - 1);
					put(&zero, 4); put(&qlen, 4);
					const std::string& id = translated ? dq.ids[sq] : q.ids[sq];
					put(id.c_str(), id.size() + 1);
					bool has_n = false;
					if (translated) for (char c : dq.dna[sq]) has_n |= c == 'N';
					const uint8_t fl = has_n ? 1 : 0;
					put(&fl, 1);
					const unsigned bits = translated ? (has_n ? 3u : 2u) : 5u;
					unsigned acc = 0, nb = 0;
					for (uint32_t p2 = 0; p2 < qlen; ++p2) {
						const unsigned v = translated ? (unsigned)(strchr("ACGTN", dq.dna[sq][p2]) - "ACGTN") : (unsigned)(q.letters[(size_t)q.limits[sq] + p2] & 31);
						acc |= v << nb; nb += bits;
						if (nb >= 8) { const uint8_t b = (uint8_t)(acc & 0xFF); put(&b, 1); nb -= 8; acc >>= 8; }
					}
					if (nb > 0) { const uint8_t b = (uint8_t)(acc & 0xFF); put(&b, 1); }
					++n_queries;
				}
				if (dict_of[x.target] == UINT32_MAX) { dict_of[x.target] = (uint32_t)dict.size(); dict.push_back(x.target); }
				int64_t qb = x.q_begin;  // Hsp::oriented_range().begin_: 0-based, on the read for translated queries (a reverse-strand alignment begins at its high end)
				unsigned rev = 0;
				if (translated) {
					const int fr = (int)(x.query % 6), off = fr % 3;
					const int64_t b_in = 3 * (int64_t)x.q_begin + off;
					qb = fr < 3 ? b_in : (int64_t)dq.len[sq] - b_in - 1;
					rev = fr < 3 ? 0u : 1u;
				}
				const uint8_t flag = (uint8_t)(len_flag((uint32_t)x.score) | (len_flag((uint32_t)qb) << 2) | (len_flag((uint32_t)x.t_begin) << 4) | (rev << 6));
				put(&dict_of[x.target], 4); put(&flag, 1);
				put_packed((uint32_t)x.score); put_packed((uint32_t)qb); put_packed((uint32_t)x.t_begin);
				// PackedTranscript as the traceback leaves it (basic/hssp.cpp:260-290, banded_swipe.h:160-182): a match is one byte with count 1, a
				// deletion / substitution one byte with its subject letter, a frameshift a substitution by letter 26 (reverse) or 27 (forward);
				// an insertion gap of n columns is pushed as counts of 63, ..., 63, n mod 63 while walking back from the alignment's end, and the
				// whole transcript is reversed afterwards: the remainder byte comes first
				const uint8_t* t = tr + x.transcript_off;
				for (uint32_t k = 0; k < x.transcript_len;) {
					const uint8_t b = t[k];
					uint8_t c;
					if (b == DMND_TR_FRAMESHIFT_FWD || b == DMND_TR_FRAMESHIFT_REV) c = (uint8_t)((3u << 6) | (b == DMND_TR_FRAMESHIFT_FWD ? 27u : 26u));
					else if ((b >> 6) == DMND_OP_DELETION || (b >> 6) == DMND_OP_SUBSTITUTION) c = b;
					else if ((b >> 6) == DMND_OP_MATCH) c = 1u;
					else {
						uint32_t e = k;
						while (e < x.transcript_len && t[e] != DMND_TR_FRAMESHIFT_FWD && t[e] != DMND_TR_FRAMESHIFT_REV && (t[e] >> 6) == DMND_OP_INSERTION) ++e;
						const uint32_t run = e - k;
						if (run % 63u) { c = (uint8_t)((1u << 6) | (run % 63u)); put(&c, 1); }
						for (uint32_t f63 = run / 63u; f63; --f63) { c = (uint8_t)((1u << 6) | 63u); put(&c, 1); }
						k = e;
						continue;
					}
					put(&c, 1);
					++k;
				}
				const uint8_t term = 0;
				put(&term, 1);
				if (i + 1 == n || m[i + 1].query / cx != sq) { const uint32_t sz = (uint32_t)(body.size() - rec_pos - 4); std::memcpy(&body[rec_pos], &sz, 4); }
			}
			{ const uint32_t zero = 0; put(&zero, 4); }
			h2.block_size[0] = body.size();
			uint64_t names = 0;
			for (uint32_t tgt : dict) {  // Block::dict_id (data/block/block.cpp:138-150): the full title, all sequence ids ("\x01" between them) or the first id
				const std::string nm = salltitles ? r.titles[tgt] : sallseqid ? all_titles(r.titles[tgt], true, "\x01") : r.ids[tgt];
				put(nm.c_str(), nm.size() + 1); names += nm.size() + 1;
			}
			for (uint32_t tgt : dict) { const uint32_t l = (uint32_t)(r.limits[tgt + 1] - r.limits[tgt] - 1); put(&l, 4); }
			uint64_t all_letters = 0;
			for (uint32_t i = 0; i < r.size(); ++i) all_letters += (uint64_t)(r.limits[i + 1] - r.limits[i] - 1);
			h2.diamond_build = daa_build; h2.db_seqs = r.size(); h2.db_seqs_used = dict.size(); h2.db_letters = all_letters; h2.query_records = n_queries;
			h2.mode = translated ? 3 : 2; h2.gap_open = 11; h2.gap_extend = 1; h2.k = 0.041; h2.lambda = 0.267; h2.evalue = o.max_evalue;
			{ std::string mn = matrix_name; for (char& c : mn) c = (char)tolower((unsigned char)c); strncpy(h2.score_matrix, mn.c_str(), sizeof h2.score_matrix - 1); }
			h2.block_size[1] = names; h2.block_size[2] = dict.size() * sizeof(uint32_t);
			h2.block_type[0] = 1; h2.block_type[1] = 2; h2.block_type[2] = 3;
			fwrite(h1, 1, sizeof h1, out);
			fwrite(&h2, 1, sizeof h2, out);
			fwrite(body.data(), 1, body.size(), out);
			n = 0;
		}
		if (xml) {
			// XMLFormat (output/xml_format.cpp:31-176): the BLAST XML of NCBI's DTD -- a header naming the first query of the block, one <Iteration>
			// per aligned query (and per query with seed hits and no alignment: DEFAULT_REPORT_UNALIGNED), one <Hit> per target with its HSP
			// (coordinates on the read, unoriented, for translated queries), the aligned letters and BLAST's midline, and the footer without a final newline
			const dmnd_params* pp = &params;
			auto esc = [](const std::string& in, std::string& out_) {  // EscapeSequences::XML (util/util.cpp:80-86)
				for (char c : in) {
					if (c == '"') out_ += "&quot;"; else if (c == '\'') out_ += "&apos;"; else if (c == '<') out_ += "&lt;"; else if (c == '>') out_ += "&gt;"; else if (c == '&') out_ += "&amp;"; else out_ += c;
				}
			};
			auto next_title = [](const std::string& t, size_t a, size_t* next) -> std::string {  // one title of a merged record: "\x01" or " >" ends it (FASTA_HEADER_SEP)
				size_t e = a;
				while (e < t.size() && t[e] != '\x01' && !(t[e] == ' ' && e + 1 < t.size() && t[e + 1] == '>')) ++e;
				*next = e >= t.size() ? std::string::npos : (t[e] == '\x01' ? e + 1 : e + 2);
				return t.substr(a, e - a);
			};
			auto accession = [](std::string t) {  // Util::Seq::get_accession (util/sequence/sequence.cpp:76-103): UniRef / gi| / xx| prefixes, |xx and .version suffixes go
				size_t i;
				if (t.compare(0, 6, "UniRef") == 0) t.erase(0, t.find('_') + 1);
				else if ((i = t.find('|')) != std::string::npos) {
					if (t.compare(0, 3, "gi|") == 0) { t.erase(0, t.find('|', i + 1) + 1); i = t.find('|'); }
					t.erase(0, i + 1);
					i = t.find('|');
					if (i != std::string::npos) t.erase(i);
				}
				i = t.rfind('.');
				if (i != std::string::npos) t.erase(i);
				return t;
			};
			const uint32_t cx = translated ? 6u : 1u;
			auto q_title = [&](uint32_t sq) -> const std::string& { return translated ? dq.titles[sq] : q.titles[sq]; };
			auto q_len = [&](uint32_t sq) -> int64_t { return translated ? (int64_t)dq.len[sq] : q.limits[sq + 1] - q.limits[sq] - 1; };
			uint64_t all_letters = 0;
			for (uint32_t i = 0; i < r.size(); ++i) all_letters += (uint64_t)(r.limits[i + 1] - r.limits[i] - 1);
			{
				size_t nx;
				std::string e0, dbl;
				const uint32_t n_src = translated ? (uint32_t)dq.ids.size() : q.size();
				if (n_src) { std::string full; esc(q_title(0), full); e0 = full.substr(0, q_title(0).find('\x01')); }
				(void)nx;
				std::ostringstream ev;
				ev << o.max_evalue;  // (a stringstream's default formatting of the double: 0.001, 1e-10, 10)
				line = "<?xml version=\"1.0\"?>\n<!DOCTYPE BlastOutput PUBLIC \"-//NCBI//NCBI BlastOutput/EN\" \"http://www.ncbi.nlm.nih.gov/dtd/NCBI_BlastOutput.dtd\">\n<BlastOutput>\n"
				       "  <BlastOutput_program>" + std::string(translated ? "blastx" : "blastp") + "</BlastOutput_program>\n  <BlastOutput_version>diamond 2.2.2</BlastOutput_version>\n"
				       "  <BlastOutput_reference>Benjamin Buchfink, Xie Chao, and Daniel Huson (2015), &quot;Fast and sensitive protein alignment using DIAMOND&quot;, Nature Methods 12:59-60.</BlastOutput_reference>\n"
				       "  <BlastOutput_db>" + df + "</BlastOutput_db>\n  <BlastOutput_query-ID>Query_1</BlastOutput_query-ID>\n  <BlastOutput_query-def>" + e0 + "</BlastOutput_query-def>\n"
				       "  <BlastOutput_query-len>" + std::to_string(n_src ? q_len(0) : 0) + "</BlastOutput_query-len>\n  <BlastOutput_param>\n    <Parameters>\n      <Parameters_matrix>" + matrix_name + "</Parameters_matrix>\n"
				       "      <Parameters_expect>" + ev.str() + "</Parameters_expect>\n      <Parameters_gap-open>11</Parameters_gap-open>\n      <Parameters_gap-extend>1</Parameters_gap-extend>\n"
				       "      <Parameters_filter>F</Parameters_filter>\n    </Parameters>\n  </BlastOutput_param>\n<BlastOutput_iterations>\n";
				if (n_src) fwrite(line.data(), 1, line.size(), out);  // (the header is printed by the first query block that has ids: none without queries)
			}
			auto intro = [&](uint32_t sq) {
				const uint32_t oid = (!translated && !q.oid.empty()) ? q.oid[sq] : sq;
				size_t nx;
				std::string t;
				esc(next_title(q_title(sq), 0, &nx), t);
				return "<Iteration>\n  <Iteration_iter-num>" + std::to_string(oid + 1) + "</Iteration_iter-num>\n  <Iteration_query-ID>Query_" + std::to_string(oid + 1) + "</Iteration_query-ID>\n  <Iteration_query-def>"
				       + t + "</Iteration_query-def>\n  <Iteration_query-len>" + std::to_string(q_len(sq)) + "</Iteration_query-len>\n<Iteration_hits>\n";
			};
			char kb[40], lb[40];
			snprintf(kb, sizeof kb, "%lf", 0.041);  // ScoreMatrix::k() / lambda() of BLOSUM62 11/1 (TextBuffer::print_d)
			snprintf(lb, sizeof lb, "%lf", 0.267);
			const std::string epilog_tail = "</Iteration_hits>\n  <Iteration_stat>\n    <Statistics>\n      <Statistics_db-num>" + std::to_string(view_mode ? view_header.db_seqs : (uint64_t)r.size()) + "</Statistics_db-num>\n      <Statistics_db-len>" + std::to_string(view_mode ? view_header.db_letters : all_letters)
			                                + "</Statistics_db-len>\n      <Statistics_hsp-len>0</Statistics_hsp-len>\n      <Statistics_eff-space>0</Statistics_eff-space>\n      <Statistics_kappa>" + kb
			                                + "</Statistics_kappa>\n      <Statistics_lambda>" + lb + "</Statistics_lambda>\n      <Statistics_entropy>0</Statistics_entropy>\n    </Statistics>\n  </Iteration_stat>\n</Iteration>\n";
			size_t nu = 0, u = 0;
			const uint32_t* unal_ids = result_unaligned(&nu);
			auto no_hits_upto = [&](uint32_t src_end) {
				for (; u < nu && unal_ids[u] / cx < src_end; ++u) {
					line = intro(unal_ids[u] / cx) + epilog_tail;
					fwrite(line.data(), 1, line.size(), out);
				}
			};
			uint32_t hit_num = 0;
			for (size_t i = 0; i < n; ++i) {
				const dmnd_match& x = m[i];
				const uint32_t sq = x.query / cx;
				const bool first = i == 0 || m[i - 1].query / cx != sq, last = i + 1 == n || m[i + 1].query / cx != sq;
				line.clear();
				if (first) { no_hits_upto(sq); line = intro(sq); hit_num = 0; }
				const uint8_t* t = tr + x.transcript_off;
				const std::string& tt = r.titles[x.target];
				size_t d0 = 0;
				while (d0 < tt.size() && !strchr(" \a\b\f\n\r\t\v\x01", tt[d0])) ++d0;  // Util::Seq::get_title_def: id | definition at the first id delimiter
				const std::string id = tt.substr(0, d0), def = d0 >= tt.size() ? std::string() : tt.substr(d0 + 1);
				if (hit_num > 0) line += "  </Hit_hsps>\n</Hit>\n";
				line += "<Hit>\n  <Hit_num>" + std::to_string(hit_num + 1) + "</Hit_num>\n  <Hit_id>";
				if (xml_blord) line += "gnl|BL_ORD_ID|" + std::to_string(r.oid.empty() ? x.target : r.oid[x.target]); else esc(id, line);
				line += "</Hit_id>\n  <Hit_def>";
				const std::string& defs = xml_blord ? tt : def;  // --xml-blord-format: the whole title line
				for (size_t a = 0, k = 0; a != std::string::npos; ++k) {  // OutputFormat::print_title(def, full, all, " &gt;")
					size_t nx;
					const std::string one = next_title(defs, a, &nx);
					if (k) line += " &gt;";
					esc(one, line);
					a = nx;
				}
				line += "</Hit_def>\n  <Hit_accession>";
				esc(no_parse_seqids ? id : accession(id), line);
				line += "</Hit_accession>\n  <Hit_len>" + std::to_string(r.limits[x.target + 1] - r.limits[x.target] - 1) + "</Hit_len>\n  <Hit_hsps>\n";
				++hit_num;
				format_double(x.bit_score, buf, sizeof buf);
				line += "    <Hsp>\n      <Hsp_num>1</Hsp_num>\n      <Hsp_bit-score>" + std::string(buf) + "</Hsp_bit-score>\n      <Hsp_score>" + std::to_string(x.score) + "</Hsp_score>\n      <Hsp_evalue>";
				if (x.evalue == 0.0) line += "0.0"; else { snprintf(buf, sizeof buf, "%.2e", x.evalue); line += buf; }
				int64_t qf = x.q_begin, qt = x.q_end;  // query_source_range: [begin, end) on the query as given
				int frame_tag = 0;
				if (translated) {
					const int fr = (int)(x.query % 6), off = fr % 3;
					const int64_t L = dq.len[sq], b_in = 3 * (int64_t)x.q_begin + off, e_in = 3 * (int64_t)x.q_end + end_frame(x) % 3;
					if (fr < 3) { qf = b_in; qt = e_in; } else { qf = L - e_in; qt = L - b_in; }
					frame_tag = fr < 3 ? fr + 1 : 2 - fr;  // Hsp::blast_query_frame
				}
				line += "</Hsp_evalue>\n      <Hsp_query-from>" + std::to_string(qf + 1) + "</Hsp_query-from>\n      <Hsp_query-to>" + std::to_string(qt) + "</Hsp_query-to>\n      <Hsp_hit-from>" + std::to_string(x.t_begin + 1)
				        + "</Hsp_hit-from>\n      <Hsp_hit-to>" + std::to_string(x.t_end) + "</Hsp_hit-to>\n      <Hsp_query-frame>" + std::to_string(frame_tag) + "</Hsp_query-frame>\n      <Hsp_hit-frame>0</Hsp_hit-frame>\n      <Hsp_identity>"
				        + std::to_string(x.identities) + "</Hsp_identity>\n      <Hsp_positive>" + std::to_string(x.positives) + "</Hsp_positive>\n      <Hsp_gaps>" + std::to_string(x.gaps) + "</Hsp_gaps>\n      <Hsp_align-len>"
				        + std::to_string(x.length) + "</Hsp_align-len>\n         <Hsp_qseq>";
				walk_query(x, t);
				std::string ql, ml, sl;
				for (uint32_t k = 0; k < x.transcript_len; ++k) {
					const int op = t[k] >> 6, sc = t[k] & 63;
					if (qat[k] < 0) { ql += qat[k] == -1 ? '\\' : '/'; sl += '-'; ml += ' '; }
					else if (op == DMND_OP_MATCH) { const char c = alphabet[qat[k]]; ql += c; ml += c; sl += c; }
					else if (op == DMND_OP_SUBSTITUTION) { const int a = qat[k]; ql += alphabet[a]; sl += alphabet[sc]; ml += pp->score[a * 32 + sc] > 0 ? '+' : ' '; }
					else if (op == DMND_OP_INSERTION) { ql += alphabet[qat[k]]; sl += '-'; ml += ' '; }
					else { ql += '-'; sl += alphabet[sc]; ml += ' '; }
				}
				line += ql + "</Hsp_qseq>\n         <Hsp_hseq>" + sl + "</Hsp_hseq>\n      <Hsp_midline>" + ml + "</Hsp_midline>\n    </Hsp>\n";
				if (last) line += "  </Hit_hsps>\n</Hit>\n" + epilog_tail;
				fwrite(line.data(), 1, line.size(), out);
			}
			no_hits_upto(UINT32_MAX);
			line = "</BlastOutput_iterations>\n</BlastOutput>";
			fwrite(line.data(), 1, line.size(), out);
			n = 0;
		}
		if (header_simple && !pairwise && !paf && !sam && !xml && !daa) {  // TabularFormat::output_header: the field keys, tab-separated
			line.clear();
			for (size_t fi = 0; fi < fields.size(); ++fi) { if (fi) line += '\t'; line += fields[fi]; }
			line += '\n';
			fwrite(line.data(), 1, line.size(), out);
		}
		if (json && !pairwise && !paf && !sam && !xml && !daa) fwrite("[", 1, 1, out);
		size_t n_unal = 0, u_next = 0;
		const uint32_t* unal_q = result_unaligned(&n_unal);
		const uint32_t ctxs = translated ? 6u : 1u;
		auto unaligned_upto = [&](uint32_t src_end) {  // --unal 1: TabularFormat::print_query_intro (output/blast_tab_format.cpp:776-788) for the queries
			if (!unal || pairwise || paf || sam || xml || daa) return;      // [.., src_end) that had seed hits and no alignment, in query order
			for (; u_next < n_unal && unal_q[u_next] / ctxs < src_end; ++u_next) {
				const uint32_t sq = unal_q[u_next] / ctxs;
				line.clear();
				for (size_t fi = 0; fi < fields.size(); ++fi) {
					const std::string& f = fields[fi];
					if (fi) line += '\t';
					if (f == "qseqid") line += translated ? dq.ids[sq] : q.ids[sq];
					else if (f == "qlen") line += std::to_string(translated ? (int64_t)dq.len[sq] : q.limits[sq + 1] - q.limits[sq] - 1);
					else if (f == "qtitle") line += translated ? dq.titles[sq] : q.titles[sq];
					else if (f == "qframe") line += '0';
					else if (f == "hspnum" || f == "normalized_nident") throw std::runtime_error("Invalid output field: " + f);  // no unaligned form in the reference either (make_invalid_intro_handler, blast_tab_format.cpp:119-124)
					else if (f == "full_qseq") { if (translated) line += dq.dna[sq]; else for (int64_t p2 = q.limits[sq]; p2 < q.limits[sq + 1] - 1; ++p2) line += alphabet[q.letters[(size_t)p2] & 31]; }
					else if (f == "full_qqual") line += (translated && sq < dq.qual.size() && !dq.qual[sq].empty()) ? dq.qual[sq] : std::string("*");
					else if (f == "sallseqid" || f == "salltitles" || f == "full_sseq" || f == "qqual" || f == "qseq_translated") line += '*';
					else if (f == "sseqid" || f == "cigar" || f == "btop" || f == "qseq_gapped" || f == "sseq_gapped" || f == "stitle" || f == "qstrand" || f == "qseq" || f == "sseq") line += '*';
					else line += "-1";
				}
				line += '\n';
				fwrite(line.data(), 1, line.size(), out);
			}
		};
		for (size_t i = 0; i < n; ++i) {
			const dmnd_match& x = m[i];
			unaligned_upto(x.query / ctxs);
			const uint8_t* t = tr + x.transcript_off;
			const int8_t* qs = q.letters.data() + q.limits[x.query];
			line.clear();
			for (size_t fi = 0; fi < fields.size(); ++fi) {
				const std::string& f = fields[fi];
				// json-flat: "key":value lines, strings quoted, the arrays (sallseqid, salltitles) bracketed with every element quoted
				const bool j_arr = json && (f == "sallseqid" || f == "salltitles");
				const bool j_str = json && (f == "qseqid" || f == "sseqid" || f == "qseq" || f == "sseq" || f == "btop" || f == "stitle" || f == "qtitle" || f == "full_sseq" || f == "qqual" || f == "full_qqual"
				                            || f == "full_qseq" || f == "qseq_gapped" || f == "sseq_gapped" || f == "qstrand" || f == "cigar" || f == "qseq_translated");
				if (json) { line += "\t\"" + f + "\":"; if (j_str) line += '"'; if (j_arr) line += "[\""; }
				else if (fi) line += '\t';
				const size_t value_at = line.size();
				if (f == "qseqid") line += translated ? dq.ids[x.query / 6] : q.ids[x.query];
				else if (f == "sseqid") line += r.ids[x.target];
				else if (f == "pident") { format_double((double)x.identities * 100.0 / (double)x.length, buf, sizeof buf); line += buf; }
				else if (f == "length") line += std::to_string(x.length);
				else if (f == "mismatch") line += std::to_string(x.mismatches);
				else if (f == "gapopen") line += std::to_string(x.gap_openings);
				else if (translated && (f == "qstart" || f == "qend")) {
					// TranslatedPosition::absolute_interval (basic/translated_position.h:121-127): in-strand = 3 * translated + frame offset;
					// a reverse-strand range is mirrored, and printed from its high end (output/blast_tab_format.cpp, query_source_range)
					const int fr = (int)(x.query % 6), off = fr % 3, L = dq.len[x.query / 6];
					const int b_in = 3 * x.q_begin + off, e_in = 3 * x.q_end + end_frame(x) % 3;  // Hsp::set_begin / set_end (basic/hssp.cpp:197-217): begin and end in their own frames
					if (fr < 3) line += std::to_string(f == "qstart" ? b_in + 1 : e_in);
					else line += std::to_string(f == "qstart" ? L - b_in : L - e_in + 1);
				}
				else if (f == "qstart") line += std::to_string(x.q_begin + 1);
				else if (f == "qend") line += std::to_string(x.q_end);
				else if (f == "sstart") line += std::to_string(x.t_begin + 1);
				else if (f == "send") line += std::to_string(x.t_end);
				else if (f == "evalue") { if (x.evalue == 0.0) line += "0.0"; else { snprintf(buf, sizeof buf, "%.2e", x.evalue); line += buf; } }
				else if (f == "bitscore") { format_double(x.bit_score, buf, sizeof buf); line += buf; }
				else if (f == "score") line += std::to_string(x.score);
				else if (f == "qseq") {  // the aligned stretch of the SOURCE query: letters of the read (forward orientation) for blastx
					if (!translated) for (int p2 = x.q_begin; p2 < x.q_end; ++p2) line += alphabet[qs[p2] & 31];
					else {
						const int fr = (int)(x.query % 6), off = fr % 3;
						const int64_t L = dq.len[x.query / 6], b_in = 3 * (int64_t)x.q_begin + off, e_in = 3 * (int64_t)x.q_end + end_frame(x) % 3;  // (the end lies in its own frame after a frameshift)
						const int64_t b = fr < 3 ? b_in : L - e_in, e = fr < 3 ? e_in : L - b_in;
						line.append(dq.dna[x.query / 6], (size_t)b, (size_t)(e - b));
					}
				}
				else if (f == "sseq") {  // the subject letters of the alignment (no gap characters)
					walk_query(x, t);  // (a match byte stands for the query letter it consumes; frameshift bytes consume nucleotides, no subject letter)
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						if (qat[k] < 0) {  // HspContext::Iterator::subject() of a frameshift operation falls through to query(): the query letter at the
							// position before the shift (basic/match.h:325-334) -- the reference's sseq carries that letter, so does this one
							const uint32_t c0 = x.query - x.query % 3;
							line += alphabet[q.letters[(size_t)q.limits[c0 + (uint32_t)(qcur[k] % 3)] + (size_t)(qcur[k] / 3)] & 31];
							continue;
						}
						const int o2 = t[k] >> 6;
						if (o2 == DMND_OP_MATCH) line += alphabet[qat[k]];
						else if (o2 != DMND_OP_INSERTION) line += alphabet[t[k] & 63];
					}
				}
				else if (f == "qtitle") line += translated ? dq.titles[x.query / 6] : q.titles[x.query];
				else if (f == "stitle") { const std::string& tt = r.titles[x.target]; line.append(tt, 0, std::min(tt.find('\x01'), tt.find(" >"))); }  // print_title(full titles, first one only: "\x01" or " >" separate merged records)
				else if (f == "sallseqid" || f == "salltitles") {  // print_title over every title of a merged record ("\x01" or " >" between them): ids joined by ';', titles by "<>"
					const std::string& tt = r.titles[x.target];
					const bool ids_only = f == "sallseqid";
					for (size_t a = 0, k = 0; a <= tt.size(); ++k) {
						size_t e = a;
						while (e < tt.size() && tt[e] != '\x01' && !(tt[e] == ' ' && e + 1 < tt.size() && tt[e + 1] == '>')) ++e;
						if (k) line += json ? "\",\"" : ids_only ? ";" : "<>";
						size_t ie = e;
						if (ids_only) { ie = a; while (ie < e && !strchr(" \a\b\f\n\r\t\v", tt[ie])) ++ie; }
						line.append(tt, a, ie - a);
						if (e >= tt.size()) break;
						a = tt[e] == '\x01' ? e + 1 : e + 2;
					}
				}
				else if (f == "full_sseq") for (int64_t p2 = r.limits[x.target]; p2 < r.limits[x.target + 1] - 1; ++p2) line += alphabet[r_unmasked[(size_t)p2] & 31];  // the target as loaded (Output::Flags::TARGET_SEQS keeps the unmasked block, run/double_indexed.cpp:118-120)
				else if (f == "full_qseq") { if (translated) line += dq.dna[x.query / 6]; else for (int64_t p2 = q.limits[x.query]; p2 < q.limits[x.query + 1] - 1; ++p2) line += alphabet[q.letters[(size_t)p2] & 31]; }
				else if (f == "qnum") line += std::to_string(translated ? x.query / 6 : (q.oid.empty() ? x.query : q.oid[x.query]));
				else if (f == "snum") line += std::to_string(r.oid.empty() ? x.target : r.oid[x.target]);
				else if (f == "hspnum") line += '0';
				else if (f == "approx_pident") {  // Hsp::approx_id_percent (basic/hssp.cpp:381-392): 100 for identical stretches, else Stats::approx_id of the score per column
					const int ql = x.q_end - x.q_begin, tl = x.t_end - x.t_begin;
					bool identical = ql == tl;
					for (int k2 = 0; identical && k2 < ql; ++k2) identical = (qs[x.q_begin + k2] & 31) == (r.letters[(size_t)r.limits[x.target] + (size_t)(x.t_begin + k2)] & 31);
					const int mx = std::max(ql, tl);
					format_double(fshift ? 0.0 /* the 3-frame DP leaves Hsp::approx_id at 0 */ : identical || mx == 0 ? 100.0 : std::min(std::max(std::fma((double)x.score / mx, 16.56, 11.41), 0.0), 100.0), buf, sizeof buf);
					line += buf;
				}
				else if (f == "normalized_nident") { snprintf(buf, sizeof buf, "%lf", (double)x.identities / (double)std::max<int64_t>(q.limits[x.query + 1] - q.limits[x.query] - 1, r.limits[x.target + 1] - r.limits[x.target] - 1)); line += buf; }
				else if (f == "qseq_translated") {  // the aligned letters of the frame; in frameshift mode the query letters the transcript consumes (blast_tab_format.cpp:565-576)
					walk_query(x, t);
					for (uint32_t k = 0; k < x.transcript_len; ++k) if (qat[k] >= 0 && (t[k] >> 6) != DMND_OP_DELETION) line += alphabet[qat[k]];
				}
				else if (f == "qqual" || f == "full_qqual") {
					const uint32_t sq = x.query / 6;
					if (!translated || sq >= dq.qual.size() || dq.qual[sq].empty()) line += '*';
					else if (f == "full_qqual") line += dq.qual[sq];
					else {
						const int fr = (int)(x.query % 6), off = fr % 3;
						const int64_t L = dq.len[sq], b_in = 3 * (int64_t)x.q_begin + off, e_in = 3 * (int64_t)x.q_end + end_frame(x) % 3;
						const int64_t b = fr < 3 ? b_in : L - e_in, e = fr < 3 ? e_in : L - b_in;
						line.append(dq.qual[sq], (size_t)b, (size_t)(e - b));
					}
				}
				else if (f == "positive") line += std::to_string(x.positives);
				else if (f == "ppos") { format_double((double)x.positives * 100.0 / (double)x.length, buf, sizeof buf); line += buf; }
				else if (f == "qcovhsp") {  // query_source_range().length() * 100 / source length (basic/match.h): nucleotides for blastx
					const double cov = translated ? (double)(3 * (x.q_end - x.q_begin) + end_frame(x) % 3 - (int)(x.query % 3)) * 100.0 / (double)dq.len[x.query / 6]
					                              : (double)(x.q_end - x.q_begin) * 100.0 / (double)(q.limits[x.query + 1] - q.limits[x.query] - 1);
					format_double(cov, buf, sizeof buf); line += buf;
				}
				else if (f == "scovhsp") { format_double((double)(x.t_end - x.t_begin) * 100.0 / (double)(r.limits[x.target + 1] - r.limits[x.target] - 1), buf, sizeof buf); line += buf; }
				else if (f == "qframe") { const int fr = (int)(x.query % 6); line += std::to_string(translated ? (fr < 3 ? fr + 1 : 2 - fr) : 0); }
				else if (f == "qstrand") line += (translated && x.query % 6 >= 3) ? '-' : '+';
				else if (f == "gaps") line += std::to_string(x.gaps);
				else if (f == "nident") line += std::to_string(x.identities);
				else if (f == "qlen") line += std::to_string(translated ? (int64_t)dq.len[x.query / 6] : q.limits[x.query + 1] - q.limits[x.query] - 1);
				else if (f == "slen") line += std::to_string(r.limits[x.target + 1] - r.limits[x.target] - 1);
				else if (f == "cigar") {  // print_cigar, output/sam_format.cpp:67-83: match and substitution are both M, frameshifts \ and /
					uint32_t run = 0; int op = -1;
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						const int o2 = t[k] >> 6, c = t[k] == DMND_TR_FRAMESHIFT_FWD ? 3 : t[k] == DMND_TR_FRAMESHIFT_REV ? 4 : (o2 == DMND_OP_INSERTION) ? 1 : (o2 == DMND_OP_DELETION) ? 2 : 0;
						if (c == op) ++run; else { if (run) { line += std::to_string(run); line += "MID\\/"[op]; } run = 1; op = c; }
					}
					if (run) { line += std::to_string(run); line += "MID\\/"[op]; }
				}
				else if (f == "btop") {  // output/blast_tab_format.cpp:365-398
					uint32_t nm = 0;
					walk_query(x, t);
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						const int o2 = t[k] >> 6;
						if (o2 == DMND_OP_MATCH) { ++nm; continue; }
						if (nm) { line += std::to_string(nm); nm = 0; }
						if (qat[k] < 0) { line += qat[k] == -1 ? '\\' : '/'; line += '-'; }
						else if (o2 == DMND_OP_SUBSTITUTION) { line += alphabet[qat[k]]; line += alphabet[t[k] & 63]; }
						else if (o2 == DMND_OP_INSERTION) { line += alphabet[qat[k]]; line += '-'; }
						else { line += '-'; line += alphabet[t[k] & 63]; }
					}
					if (nm) line += std::to_string(nm);
				}
				else if (f == "qseq_gapped" || f == "sseq_gapped") {
					const bool query_side = f[0] == 'q';
					walk_query(x, t);
					for (uint32_t k = 0; k < x.transcript_len; ++k) {
						const int o2 = t[k] >> 6;
						if (qat[k] < 0) { line += query_side ? (qat[k] == -1 ? '\\' : '/') : '-'; continue; }  // HspContext::Iterator::query_char / subject_char
						const char qc = (o2 == DMND_OP_DELETION) ? '-' : alphabet[qat[k]];
						const char sc = (o2 == DMND_OP_INSERTION) ? '-' : (o2 == DMND_OP_MATCH ? alphabet[qat[k]] : alphabet[t[k] & 63]);
						line += query_side ? qc : sc;
					}
				}
				(void)value_at;
				if (json) { if (j_str) line += '"'; if (j_arr) line += "\"]"; line += fi + 1 < fields.size() ? ",\n" : "\n"; }
			}
			if (json) line = std::string(i ? "," : "") + "\n\t{\n" + line + "\t}";  // (a comma between the records of one query, the query separator between queries: one after every record but the last)
			else line += '\n';
			fwrite(line.data(), 1, line.size(), out);
		}
		unaligned_upto(UINT32_MAX);
		if (json) fwrite("\n]", 1, 2, out);
		if (out == stdout) fflush(out); else fclose(out);
		if (gz_out) {  // --compress 1: the output as a gzip file, ".gz" appended to its name (basic/config.cpp:770-771)
			const std::string gzname = (no_auto_append || (of.size() >= 3 && of.compare(of.size() - 3, 3, ".gz") == 0)) ? of : of + ".gz";
			std::ifstream in(of, std::ios::binary);
			std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
			in.close();
			const std::string tmp = gzname + ".tmp";
			gzFile g = gzopen(tmp.c_str(), "wb");
			if (!g) throw std::runtime_error("Error opening file " + gzname);
			for (size_t off = 0; off < data.size();) { const unsigned chunk = (unsigned)std::min<size_t>(data.size() - off, 1u << 30); if (gzwrite(g, data.data() + off, chunk) != (int)chunk) { gzclose(g); throw std::runtime_error("Error writing file " + gzname); } off += chunk; }
			gzclose(g);
			std::remove(of.c_str());
			std::rename(tmp.c_str(), gzname.c_str());
		}
		if (log && !view_mode) {
			const dmnd_run_stats* s = dmnd_result_stats(res);
			fprintf(stderr, "Seed partition bits = %d\n", params.seedp_bits);
			fprintf(stderr, "Seeds hit             = %llu\n", (unsigned long long)s->seed.seeds_hit);
			fprintf(stderr, "Hits (filter stage 0) = %llu\n", (unsigned long long)s->seed.seed_hits);
			fprintf(stderr, "Hits (filter stage 1) = %llu\n", (unsigned long long)s->seed.tentative_matches1);
			fprintf(stderr, "Hits (filter stage 2) = %llu\n", (unsigned long long)s->seed.tentative_matches2);
			fprintf(stderr, "Hits (filter stage 3) = %llu\n", (unsigned long long)s->seed.tentative_matches3);
			fprintf(stderr, "Target hits (stage 0) = %llu\n", (unsigned long long)s->targets);
			fprintf(stderr, "Target hits (stage 3) = %llu\n", (unsigned long long)s->targets_extended);
			fprintf(stderr, "DP problems round 1/2 = %llu / %llu\n", (unsigned long long)s->dp_problems_round1, (unsigned long long)s->dp_problems_round2);
			fprintf(stderr, "DP cells round 1/2    = %llu / %llu\n", (unsigned long long)s->cells_round1, (unsigned long long)s->cells_round2);
			fprintf(stderr, "Time seed/bridge/dp1/dp2/total (ms) = %.2f / %.2f / %.2f / %.2f / %.2f\n", s->seed_ms, s->host_bridge_ms, s->dp1_ms, s->dp2_ms, s->total_ms);
			fprintf(stderr, "%llu queries aligned.\n", (unsigned long long)s->queries_aligned);
		}
		for (dmnd_result* rr : results) dmnd_result_free(rr);
		dmnd_destroy(ctx);
		return 0;
	}
	catch (const std::exception& e) {
		fprintf(stderr, "Error: %s\n", e.what());
		return 1;
	}
}
