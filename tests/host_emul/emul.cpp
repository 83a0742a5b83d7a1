// TEST INFRASTRUCTURE.  Compiles kwage_b200/csrc/bitops.cuh as plain host C++ (the CUDA intrinsics
// are shimmed in that header) and exposes the device arithmetic to the CPU test-suite, so that
// encode / window extraction / canonical / murmur3 / bit-matrix transpose / bit-sliced counters can
// be checked bit-for-bit against the oracle without a GPU.  The loops below mirror how the kernels
// drive those helpers (kmer_scan_kernel stages 1-2, transpose_kernel, search_count_kernel).
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../kwage_b200/csrc/bitops.cuh"

using namespace kwg;

extern "C" {

// Mirrors kmer_scan_kernel: encode a batch tile by tile, mark read starts, and emit the canonical
// word + `nh` hashes of every valid window in stream order.  Returns the number of k-mers.
uint64_t emu_scan(const char* bases, uint64_t n_bases, const uint64_t* offsets, uint64_t n_reads, uint32_t k,
	uint32_t nh, uint64_t* out_words, uint64_t* out_pos, uint32_t* out_hash)
{
	const uint32_t TILE_BASES = 4096, TILE_LOAD = 4096 + 32, TILE_VEC = TILE_LOAD / 16;
	std::vector<uint32_t> start_mask(n_bases / 32 + 2, 0);
	for (uint64_t r = 0; r < n_reads; ++r) {
		const uint64_t p = offsets[r] - offsets[0];
		if (p < n_bases) start_mask[p >> 5] |= 1u << (p & 31);
	}
	uint64_t n = 0;
	for (uint64_t t0 = 0; t0 < n_bases; t0 += TILE_BASES) {
		uint32_t s_codes[TILE_VEC + 2], s_bad[TILE_LOAD / 32 + 2], s_start[TILE_LOAD / 32 + 2];
		uint16_t* bad16p = reinterpret_cast<uint16_t*>(s_bad);
		for (uint32_t v = 0; v < TILE_VEC; ++v) {
			const uint64_t g = t0 + (uint64_t)v * 16;
			uint32_t codes = 0, bad16 = 0xFFFFu;
			if (g < n_bases) {
				uint32_t w[4] = {0, 0, 0, 0};
				for (uint32_t j = 0; j < 16; ++j) {
					const uint32_t b = (g + j < n_bases) ? (uint8_t)bases[g + j] : (uint32_t)'N';
					w[j >> 2] |= b << (8 * (j & 3));
				}
				encode16(make_uint4(w[0], w[1], w[2], w[3]), codes, bad16);
			}
			s_codes[v] = codes;
			bad16p[v] = (uint16_t)bad16;
		}
		for (uint32_t v = 0; v < TILE_LOAD / 32 + 1; ++v) {
			const uint64_t w = (t0 >> 5) + v;
			s_start[v] = (w * 32 < n_bases) ? start_mask[w] : 0u;
		}
		s_codes[TILE_VEC] = 0; s_codes[TILE_VEC + 1] = 0;
		s_bad[TILE_LOAD / 32] = 0xFFFFFFFFu; s_bad[TILE_LOAD / 32 + 1] = 0xFFFFFFFFu;
		s_start[TILE_LOAD / 32 + 1] = 0;

		for (uint32_t p = 0; p < TILE_BASES; ++p) {
			if (!window_ok(s_bad, s_start, p, k)) continue;
			const Canon c = canonical(window_sense(s_codes, p, k), k);
			out_words[n] = c.word;
			out_pos[n] = t0 + p;
			uint32_t h[8];
			switch (nh) {
				case 1: { uint32_t x[1]; murmur3_multi<1>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 2: { uint32_t x[2]; murmur3_multi<2>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 3: { uint32_t x[3]; murmur3_multi<3>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 4: { uint32_t x[4]; murmur3_multi<4>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 5: { uint32_t x[5]; murmur3_multi<5>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 6: { uint32_t x[6]; murmur3_multi<6>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				case 7: { uint32_t x[7]; murmur3_multi<7>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
				default: { uint32_t x[8]; murmur3_multi<8>(c.low, k, x); memcpy(h, x, sizeof(x)); break; }
			}
			for (uint32_t s = 0; s < nh; ++s) out_hash[n * nh + s] = h[s];
			++n;
		}
	}
	return n;
}

// Mirrors kmer_scan_wide_kernel (raw mode, k in 1..63 through the 128-bit helpers): sets bit (hash_h & mask), h < nh, of every
// valid window.  Returns the number of k-mers.
uint64_t emu_scan_wide_insert(const char* bases, uint64_t n_bases, const uint64_t* offsets, uint64_t n_reads, uint32_t k,
	uint32_t nh, uint32_t log2_len, uint8_t* bits)
{
	const uint32_t TILE_BASES = 4096, TILE_LOAD = 4096 + 64, TILE_VEC = TILE_LOAD / 16;
	const uint32_t mask = (log2_len >= 32) ? 0xFFFFFFFFu : ((1u << log2_len) - 1u);
	std::vector<uint32_t> start_mask(n_bases / 32 + 4, 0);
	for (uint64_t r = 0; r < n_reads; ++r) {
		const uint64_t p = offsets[r] - offsets[0];
		if (p < n_bases) start_mask[p >> 5] |= 1u << (p & 31);
	}
	uint64_t n = 0;
	for (uint64_t t0 = 0; t0 < n_bases; t0 += TILE_BASES) {
		uint32_t s_codes[TILE_VEC + 4], s_bad[TILE_LOAD / 32 + 3], s_start[TILE_LOAD / 32 + 3];
		uint16_t* bad16p = reinterpret_cast<uint16_t*>(s_bad);
		for (uint32_t v = 0; v < TILE_VEC; ++v) {
			const uint64_t g = t0 + (uint64_t)v * 16;
			uint32_t codes = 0, bad16 = 0xFFFFu;
			if (g < n_bases) {
				uint32_t w[4] = {0, 0, 0, 0};
				for (uint32_t j = 0; j < 16; ++j) {
					const uint32_t b = (g + j < n_bases) ? (uint8_t)bases[g + j] : (uint32_t)'N';
					w[j >> 2] |= b << (8 * (j & 3));
				}
				encode16(make_uint4(w[0], w[1], w[2], w[3]), codes, bad16);
			}
			s_codes[v] = codes;
			bad16p[v] = (uint16_t)bad16;
		}
		for (uint32_t v = 0; v < TILE_LOAD / 32 + 1; ++v) {
			const uint64_t w = (t0 >> 5) + v;
			s_start[v] = (w * 32 < n_bases) ? start_mask[w] : 0u;
		}
		for (uint32_t v = TILE_VEC; v < TILE_VEC + 4; ++v) s_codes[v] = 0;
		s_bad[TILE_LOAD / 32] = 0xFFFFFFFFu; s_bad[TILE_LOAD / 32 + 1] = 0xFFFFFFFFu; s_bad[TILE_LOAD / 32 + 2] = 0xFFFFFFFFu;
		s_start[TILE_LOAD / 32 + 1] = 0; s_start[TILE_LOAD / 32 + 2] = 0;
		for (uint32_t p = 0; p < TILE_BASES; ++p) {
			if (!window_ok_wide(s_bad, s_start, p, k)) continue;
			const CanonWide c = canonical_wide(window_sense_wide(s_codes, p, k), k);
			uint32_t h[8];
			murmur3_multi_wide<8>(c.low, k, h);
			for (uint32_t s = 0; s < nh; ++s) { const uint32_t b = h[s] & mask; bits[b >> 3] |= (uint8_t)(1u << (b & 7)); }
			++n;
		}
	}
	return n;
}

// window_ok_word against window_ok over a bitmap pair: returns the number of positions where they differ
uint32_t emu_window_ok_word_check(const uint32_t* bad, const uint32_t* start, uint32_t n_words, uint32_t k)
{
	uint32_t diff = 0;
	for (uint32_t v = 0; v + 2 < n_words; ++v) {
		const uint32_t w = window_ok_word(bad, start, v, k);
		for (uint32_t i = 0; i < 32; ++i) diff += (uint32_t)(((w >> i) & 1u) != (window_ok(bad, start, 32 * v + i, k) ? 1u : 0u));
	}
	return diff;
}

// hash of a word given in the reference layout (as insert_words_kernel / query_kmers_kernel do)
void emu_hash_word(uint64_t word, uint32_t k, uint32_t* out5)
{
	uint32_t h[5];
	murmur3_multi<5>(reverse_groups(word, k), k, h);
	memcpy(out5, h, sizeof(h));
}

void emu_transpose32(const uint32_t* in, uint32_t* out)
{
	uint32_t a[32];
	memcpy(a, in, sizeof(a));
	transpose32(a);
	memcpy(out, a, sizeof(a));
}

// Mirrors search_count_kernel for one 128-column lane: `n` AND-ed match vectors (uint4 each) are
// dealt round-robin to `nsub` substreams, each accumulates with the Harley-Seal block of 16, then
// the substreams are merged bit-sliced and expanded.  counts: 128 uint32.
void emu_count128(const uint32_t* vecs_all /* n x 4 */, uint32_t n_all, uint32_t nsub, uint32_t* counts)
{
	// search_count_kernel: segments of search_seg_cap(nsub) k-mers, each counted by nsub substreams of 4 + 6 planes,
	// merged bit-sliced and added to the counts of the segments before it
	const int LOW = 4, UP = 6, PL = LOW + UP;
	const uint32_t seg_cap = search_seg_cap(nsub);
	for (int i = 0; i < 128; ++i) counts[i] = 0;
	for (uint32_t seg0 = 0; seg0 == 0 || seg0 < n_all; seg0 += seg_cap) {
		const uint32_t n = (n_all > seg0) ? std::min(seg_cap, n_all - seg0) : 0u;
		const uint32_t* vecs = vecs_all + (size_t)seg0 * 4;
		std::vector<uint32_t> planes((size_t)nsub * PL * 4, 0);
		const uint32_t n_blk = (n + 16 * nsub - 1) / (16 * nsub);
		for (uint32_t sub = 0; sub < nsub; ++sub) {
			uint4 pl[PL];
			for (int i = 0; i < PL; ++i) pl[i] = make_uint4(0, 0, 0, 0);
			for (uint32_t blk = 0; blk < n_blk; ++blk) {
				uint4 fA = make_uint4(0, 0, 0, 0), eA = make_uint4(0, 0, 0, 0);
				for (int quad = 0; quad < 4; ++quad) {
					uint4 v[4];
					for (int u = 0; u < 4; ++u) {
						const uint32_t i = (blk * 16 + quad * 4 + u) * nsub + sub;
						v[u] = make_uint4(0, 0, 0, 0);
						if (i < n) v[u] = make_uint4(vecs[4 * i], vecs[4 * i + 1], vecs[4 * i + 2], vecs[4 * i + 3]);
					}
					uint4 tA, tB, f;
					csa(pl[0], tA, v[0], v[1]);
					csa(pl[0], tB, v[2], v[3]);
					csa(pl[1], f, tA, tB);
					if (quad == 0 || quad == 2) fA = f;
					else {
						uint4 e;
						csa(pl[2], e, fA, f);
						if (quad == 1) eA = e;
						else {
							uint4 c16;
							csa(pl[3], c16, eA, e);
							for (int up = LOW; up < PL; ++up) {
								const uint4 t = and4(pl[up], c16);
								pl[up].x ^= c16.x; pl[up].y ^= c16.y; pl[up].z ^= c16.z; pl[up].w ^= c16.w;
								c16 = t;
							}
						}
					}
				}
			}
			for (int i = 0; i < PL; ++i) {
				uint32_t* d = &planes[((size_t)sub * PL + i) * 4];
				d[0] = pl[i].x; d[1] = pl[i].y; d[2] = pl[i].z; d[3] = pl[i].w;
			}
		}
		for (uint32_t w = 0; w < 4; ++w) {
			uint32_t tot[16];
			for (int i = 0; i < 16; ++i) tot[i] = 0;
			for (uint32_t s = 0; s < nsub; ++s) {
				uint32_t x[PL];
				for (int i = 0; i < PL; ++i) x[i] = planes[((size_t)s * PL + i) * 4 + w];
				bitsliced_add<PL>(tot, x);
			}
			for (int nb = 0; nb < 8; ++nb) {
				const uint4 c = expand_counts4(tot, nb);
				uint32_t* o = counts + w * 32 + nb * 4;
				o[0] += c.x; o[1] += c.y; o[2] += c.z; o[3] += c.w;
			}
		}
	}
}

// Mirrors the early exit of search_count_kernel<NH, true> for a chunk of `n_lanes` 128-column lanes (all of them active):
// the substreams run as independent workers that take turns in the order `order` (entry = substream; every entry lets
// that substream do its next block of 16), publish `exact maximum of my partial counts over the chunk's columns + k-mers I
// have not looked at` after the blocks search_exit_check_at() names, read the bounds the others published last (possibly
// stale), and stop once the sum is below `need`.  Returns the number of substreams that stopped early; counts (n_lanes x
// 128) are what the kernel would hand to hits_kernel: full counts when nobody stopped, partial ones otherwise.
uint32_t emu_count_exit(const uint32_t* vecs /* n x n_lanes x 4 */, uint32_t n, uint32_t n_lanes, uint32_t nsub, uint32_t need,
	const uint32_t* order, uint32_t n_order, uint32_t* counts)
{
	const int LOW = 4, UP = 6, PL = LOW + UP;
	const uint32_t n_blk = (n + 16 * nsub - 1) / (16 * nsub);
	const uint32_t blk_first = search_exit_first_blk(n, need, nsub);
	std::vector<uint4> pl((size_t)nsub * n_lanes * PL, make_uint4(0, 0, 0, 0));
	std::vector<uint32_t> ub(nsub), next_blk(nsub, 0), stopped(nsub, 0);
	for (uint32_t s = 0; s < nsub; ++s) ub[s] = search_sub_total(n, s, nsub);
	uint32_t n_stopped = 0;
	auto step = [&](uint32_t sub) {
		if (stopped[sub] || next_blk[sub] >= n_blk) return;
		const uint32_t blk = next_blk[sub]++;
		for (uint32_t lane = 0; lane < n_lanes; ++lane) {
			uint4* P = &pl[((size_t)sub * n_lanes + lane) * PL];
			uint4 fA = make_uint4(0, 0, 0, 0), eA = make_uint4(0, 0, 0, 0);
			for (int quad = 0; quad < 4; ++quad) {
				uint4 v[4];
				for (int u = 0; u < 4; ++u) {
					const uint32_t i = (blk * 16 + quad * 4 + u) * nsub + sub;
					v[u] = make_uint4(0, 0, 0, 0);
					if (i < n) { const uint32_t* q = vecs + ((size_t)i * n_lanes + lane) * 4; v[u] = make_uint4(q[0], q[1], q[2], q[3]); }
				}
				uint4 tA, tB, f;
				csa(P[0], tA, v[0], v[1]);
				csa(P[0], tB, v[2], v[3]);
				csa(P[1], f, tA, tB);
				if (quad == 0 || quad == 2) fA = f;
				else {
					uint4 e;
					csa(P[2], e, fA, f);
					if (quad == 1) eA = e;
					else {
						uint4 c16;
						csa(P[3], c16, eA, e);
						for (int up = LOW; up < PL; ++up) {
							const uint4 t = and4(P[up], c16);
							P[up].x ^= c16.x; P[up].y ^= c16.y; P[up].z ^= c16.z; P[up].w ^= c16.w;
							c16 = t;
						}
					}
				}
			}
		}
		if (!search_exit_check_at(blk, blk_first, n_blk)) return;
		// exact maximum over the chunk's columns: descent through the planes, the lanes of the substream decide together
		std::vector<uint4> cand(n_lanes, make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
		uint32_t mx = 0;
		for (int p = PL - 1; p >= 0; --p) {
			bool some = false;
			std::vector<uint4> t(n_lanes);
			for (uint32_t lane = 0; lane < n_lanes; ++lane) {
				t[lane] = and4(cand[lane], pl[((size_t)sub * n_lanes + lane) * PL + p]);
				some = some || (t[lane].x | t[lane].y | t[lane].z | t[lane].w) != 0u;
			}
			if (some) { cand = t; mx |= 1u << p; }
		}
		ub[sub] = mx + search_sub_left(search_sub_total(n, sub, nsub), blk);
		uint32_t left = 0;
		for (uint32_t s2 = 0; s2 < nsub; ++s2) left += ub[s2];
		if (left < need) { stopped[sub] = 1; ++n_stopped; }
	};
	for (uint32_t o = 0; o < n_order; ++o) step(order[o] % nsub);
	for (uint32_t sub = 0; sub < nsub; ++sub) while (!stopped[sub] && next_blk[sub] < n_blk) step(sub);   // whoever is not done finishes
	for (uint32_t lane = 0; lane < n_lanes; ++lane)
		for (uint32_t w = 0; w < 4; ++w) {
			uint32_t tot[16];
			for (int i = 0; i < 16; ++i) tot[i] = 0;
			for (uint32_t s2 = 0; s2 < nsub; ++s2) {
				uint32_t x[PL];
				for (int i = 0; i < PL; ++i) {
					const uint4 q = pl[((size_t)s2 * n_lanes + lane) * PL + i];
					x[i] = w == 0 ? q.x : w == 1 ? q.y : w == 2 ? q.z : q.w;
				}
				bitsliced_add<PL>(tot, x);
			}
			for (int nb = 0; nb < 8; ++nb) {
				const uint4 c = expand_counts4(tot, nb);
				uint32_t* o = counts + (size_t)lane * 128 + w * 32 + nb * 4;
				o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
			}
		}
	return n_stopped;
}

uint32_t emu_seg_cap(uint32_t nsub) { return search_seg_cap(nsub); }

uint64_t emu_synth_rnd(uint64_t seed, uint64_t stream, uint64_t ctr) { return synth_rnd(seed, stream, ctr); }

} // extern "C"

// ---------------------------------------------------------------------------------------------- crc32
// Mirrors crc32.cu on the host with the SAME tables (kwage_b200/csrc/crc_tables.h): tiles counted from the end of the
// message (zero-prefixed when they stick out before byte 0), the running value xor-ed into the first word, a raw register per
// 256-byte thread chunk taken as four interleaved 64-byte chains with the slice-by-4 tables (crc_tile256_kernel) or per
// 64-byte chunk (crc_tile_kernel), and the pairwise merge raw(A||B) = x^(8|B|) * raw(A) xor raw(B) with the byte-indexed
// shift tables up to a tile and the level matrices above (crc_combine_kernel, 1024 registers per step).
#include "../../kwage_b200/csrc/crc_tables.h"

namespace {

struct CrcTables {
	std::vector<uint32_t> h;
	CrcTables() { kwg::crc32_build_tables(h); }
	uint32_t step(uint32_t c, uint32_t w) const
	{
		const uint32_t x = c ^ w;
		return h[768 + (x & 255u)] ^ h[512 + ((x >> 8) & 255u)] ^ h[256 + ((x >> 16) & 255u)] ^ h[x >> 24];
	}
	uint32_t shift_table(int l, uint32_t v) const
	{
		const uint32_t* t = &h[kwg::CRC_SHIFT_OFFSET + (size_t)l * 1024];
		return t[v & 255u] ^ t[256 + ((v >> 8) & 255u)] ^ t[512 + ((v >> 16) & 255u)] ^ t[768 + (v >> 24)];
	}
	uint32_t shift_matrix(int l, uint32_t v) const { return kwg::crc_gf2_times_host(&h[4 * 256 + (size_t)l * 32], v); }
};

const CrcTables& crc_tables()
{
	static CrcTables t;
	return t;
}

} // namespace

extern "C" {

// n_bytes % 4 == 0, n_bytes >= 4.  fast != 0: the 8 KiB-per-warp tiling of crc_tile256_kernel; else 16 KiB tiles of 64-byte chunks.
uint32_t emu_crc32(const uint8_t* data, uint64_t n_bytes, uint32_t crc_in, int fast)
{
	const CrcTables& T = crc_tables();
	const uint64_t W = n_bytes / 4;
	auto word = [&](long long w) -> uint32_t {
		if (w < 0) return 0u;
		uint32_t v;
		memcpy(&v, data + (uint64_t)w * 4, 4);
		return w == 0 ? v ^ ~crc_in : v;
	};
	const int tile_log2 = fast ? 7 : 8;                         // tile = 64 bytes * 2^tile_log2
	const uint64_t tile_words = (uint64_t)16 << tile_log2;
	const uint64_t n_tiles = (W + tile_words - 1) / tile_words;
	std::vector<uint32_t> regs(n_tiles);                        // element 0 = the END of the message
	for (uint64_t t = 0; t < n_tiles; ++t) {
		const long long start = (long long)W - (long long)(t + 1) * (long long)tile_words;
		// registers of the 64-byte chunks of the tile, forward order
		std::vector<uint32_t> c(tile_words / 16);
		for (size_t i = 0; i < c.size(); ++i) {
			uint32_t r = 0;
			for (int j = 0; j < 16; ++j) r = T.step(r, word(start + (long long)i * 16 + j));
			c[i] = r;
		}
		// merge neighbours level by level: the earlier half is shifted past the later one
		for (int l = 0; l < tile_log2; ++l) {
			std::vector<uint32_t> nx(c.size() / 2);
			for (size_t i = 0; i < nx.size(); ++i)
				nx[i] = (l < kwg::CRC2_SHIFT_LEVELS && fast ? T.shift_table(l, c[2 * i]) : T.shift_matrix(l, c[2 * i])) ^ c[2 * i + 1];
			c.swap(nx);
		}
		regs[t] = c[0];
	}
	// crc_combine_kernel: groups of 1024 registers, ten levels per step; element i + stride lies earlier in the message
	int level = tile_log2;
	while (regs.size() > 1) {
		std::vector<uint32_t> out((regs.size() + 1023) / 1024);
		for (size_t b = 0; b < out.size(); ++b) {
			uint32_t s[1024];
			for (size_t i = 0; i < 1024; ++i) s[i] = (b * 1024 + i < regs.size()) ? regs[b * 1024 + i] : 0u;
			for (int l = 0; l < 10; ++l) {
				const size_t stride = (size_t)1 << l;
				for (size_t i = 0; i < 1024; i += 2 * stride) s[i] ^= T.shift_matrix(level + l, s[i + stride]);
			}
			out[b] = s[0];
		}
		regs.swap(out);
		level += 10;
	}
	return ~regs[0];
}

// the byte-indexed shift tables are the matrices: table_l(v) == M_l * v
int emu_crc_shift_tables_agree(uint32_t v)
{
	const CrcTables& T = crc_tables();
	for (int l = 0; l < kwg::CRC2_SHIFT_LEVELS; ++l)
		if (T.shift_table(l, v) != T.shift_matrix(l, v)) return 0;
	return 1;
}

} // extern "C"
