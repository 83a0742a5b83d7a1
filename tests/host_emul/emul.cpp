// TEST INFRASTRUCTURE.  Compiles kwage_b200/csrc/bitops.cuh as plain host C++ (the CUDA intrinsics
// are shimmed in that header) and exposes the device arithmetic to the CPU test-suite, so that
// encode / window extraction / canonical / murmur3 / bit-matrix transpose / bit-sliced counters can
// be checked bit-for-bit against the oracle without a GPU.  The loops below mirror how the kernels
// drive those helpers (kmer_scan_kernel stages 1-2, transpose_kernel, search_count_kernel).
#include <cstring>
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
void emu_count128(const uint32_t* vecs /* n x 4 */, uint32_t n, uint32_t nsub, uint32_t* counts)
{
	const int LOW = 4, UP = 6, PL = LOW + UP;
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
			o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w;
		}
	}
}

uint64_t emu_synth_rnd(uint64_t seed, uint64_t stream, uint64_t ctr) { return synth_rnd(seed, stream, ctr); }

} // extern "C"
