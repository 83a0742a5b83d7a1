// Host-side tables of the device CRC-32 (crc32.cu): plain C++, also compiled into the CPU test-suite
// (tests/host_emul/emul.cpp) so that the table construction and the combine algebra are checked against zlib
// without a GPU.
//
//   [0, 1024)                        four 256-entry slice-by-4 tables of the reflected polynomial 0xEDB88320
//   [1024, +(CRC_LEVELS + 10) * 32)  32x32 GF(2) matrices: level l multiplies a raw register by x^(8 * 64 * 2^l)
//   [CRC_SHIFT_OFFSET, ...)          byte-indexed forms of the first CRC2_SHIFT_LEVELS matrices: [l][k][b] = M_l * (b << 8k)
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace kwg {

constexpr int CRC_LEVELS = 44;
constexpr int CRC2_SHIFT_LEVELS = 7;
constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr size_t CRC_SHIFT_OFFSET = (size_t)4 * 256 + (size_t)(CRC_LEVELS + 10) * 32;

inline uint32_t crc_gf2_times_host(const uint32_t* mat, uint32_t vec)
{
	uint32_t r = 0;
	for (int i = 0; vec; ++i, vec >>= 1) if (vec & 1u) r ^= mat[i];
	return r;
}

inline void crc_gf2_square(uint32_t* sq, const uint32_t* mat)
{
	for (int n = 0; n < 32; ++n) sq[n] = crc_gf2_times_host(mat, mat[n]);
}

inline void crc32_build_tables(std::vector<uint32_t>& h)
{
	h.assign(CRC_SHIFT_OFFSET + (size_t)CRC2_SHIFT_LEVELS * 4 * 256, 0u);
	for (uint32_t i = 0; i < 256; ++i) {
		uint32_t c = i;
		for (int k = 0; k < 8; ++k) c = (c & 1u) ? (CRC_POLY ^ (c >> 1)) : (c >> 1);
		h[i] = c;
	}
	for (uint32_t i = 0; i < 256; ++i)
		for (int t = 1; t < 4; ++t) h[t * 256 + i] = (h[(t - 1) * 256 + i] >> 8) ^ h[h[(t - 1) * 256 + i] & 255u];
	// operator for one zero bit, squared up to one byte (3x), then to 64 bytes (6x): level 0
	uint32_t a[32], b[32];
	a[0] = CRC_POLY;
	for (int n = 1; n < 32; ++n) a[n] = 1u << (n - 1);
	for (int q = 0; q < 9; ++q) { crc_gf2_square(b, a); for (int n = 0; n < 32; ++n) a[n] = b[n]; }
	for (int l = 0; l < CRC_LEVELS + 10; ++l) {
		for (int n = 0; n < 32; ++n) h[4 * 256 + l * 32 + n] = a[n];
		crc_gf2_square(b, a);
		for (int n = 0; n < 32; ++n) a[n] = b[n];
	}
	for (int l = 0; l < CRC2_SHIFT_LEVELS; ++l)
		for (int k = 0; k < 4; ++k)
			for (uint32_t bb = 0; bb < 256; ++bb)
				h[CRC_SHIFT_OFFSET + ((size_t)l * 4 + k) * 256 + bb] = crc_gf2_times_host(&h[4 * 256 + l * 32], bb << (8 * k));
}

} // namespace kwg
