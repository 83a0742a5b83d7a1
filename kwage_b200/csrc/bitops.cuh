// Pure bit-manipulation building blocks of the kernels.  Everything here is arithmetic on
// registers, so the same source also compiles as plain host C++ (tests/host_emul/) where the few
// CUDA intrinsics are replaced by shims: that lets the CPU test-suite check the device arithmetic
// bit-for-bit against the oracle without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define KWG_DEV __device__ __forceinline__
#else
// ---- host shims (test builds only) ----
#define KWG_DEV static inline
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint64_t __brevll(uint64_t v)
{
	uint64_t r = 0;
	for (int i = 0; i < 64; ++i) r |= ((v >> i) & 1ull) << (63 - i);
	return r;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s)
{
	s &= 31;
	return s ? ((hi << s) | (lo >> (32 - s))) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s)
{
	s &= 31;
	return s ? ((lo >> s) | (hi << (32 - s))) : lo;
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
	const uint64_t src = ((uint64_t)y << 32) | x;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		const uint32_t sel = (s >> (4 * i)) & 0xF;
		uint32_t byte = (uint32_t)(src >> (8 * (sel & 7))) & 0xFF;
		if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
#endif

namespace kwg {

// ---------------------------------------------------------------- 2-bit words
// Word layout follows reference word.h:73-104: the first (5') base of a k-mer occupies the most
// significant of its 2k bits; A=0 C=1 G=2 T=3 (word.h:19).
KWG_DEV uint64_t kmer_mask(uint32_t k)
{
	return (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
}

// Reverse the order of the k 2-bit groups of w (result: first base in the LOWEST two bits).
KWG_DEV uint64_t reverse_groups(uint64_t w, uint32_t k)
{
	uint64_t r = __brevll(w) >> (64 - 2 * k);                       // reverses bits inside each pair too
	return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}

// Canonical k-mer of a sense word (reference word.h:163-165) together with its "first base low"
// form that the hash consumes.  anti = reverse complement = ~reverse_groups(sense).
struct Canon {
	uint64_t word;   // min(sense, antisense), reference layout
	uint64_t low;    // same k-mer with base i at bits [2i, 2i+1]
};

KWG_DEV Canon canonical(uint64_t sense, uint32_t k)
{
	const uint64_t m = kmer_mask(k);
	const uint64_t rev = reverse_groups(sense, k);
	const uint64_t anti = ~rev & m;
	Canon c;
	if (sense <= anti) { c.word = sense; c.low = rev; }
	else               { c.word = anti;  c.low = ~sense & m; }   // reverse_groups(anti) == ~sense
	return c;
}

// ---------------------------------------------------------------- hash
// MurmurHash3_x86_32 over the k ASCII bytes of the k-mer, seeds 0..NH-1 evaluated together: the
// key-dependent k1 of every 4-byte block is computed once and shared by all seeds, like the
// reference's AVX2 path (hash.cpp:239-332).  `low` = k-mer with base i at bits [2i,2i+1].
KWG_DEV uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// 4 bases (8 bits, first base lowest) -> their 4 ASCII bytes, first base in byte 0.
KWG_DEV uint32_t ascii4(uint32_t c)
{
	uint32_t s = (c | (c << 4)) & 0x0F0Fu;
	s = (s | (s << 2)) & 0x3333u;                 // nibble i = code of base i
	return __byte_perm(0x54474341u, 0u, s);      // table bytes: 'A','C','G','T'
}

template <int NH>
KWG_DEV void murmur3_multi(uint64_t low, uint32_t k, uint32_t (&h)[NH], uint32_t seed0 = 0)
{
	const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
#pragma unroll
	for (int s = 0; s < NH; ++s) h[s] = seed0 + (uint32_t)s;      // seeds seed0 .. seed0 + NH - 1
	const uint32_t nblocks = k >> 2;
	for (uint32_t i = 0; i < nblocks; ++i) {
		uint32_t k1 = ascii4((uint32_t)low & 0xFFu);
		low >>= 8;
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
#pragma unroll
		for (int s = 0; s < NH; ++s) {
			uint32_t x = h[s] ^ k1;
			x = rotl32(x, 13);
			h[s] = x * 5u + 0xe6546b64u;
		}
	}
	const uint32_t rem = k & 3u;
	if (rem) {
		uint32_t k1 = ascii4((uint32_t)low & 0xFFu) & ((1u << (8 * rem)) - 1u);
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
#pragma unroll
		for (int s = 0; s < NH; ++s) h[s] ^= k1;
	}
#pragma unroll
	for (int s = 0; s < NH; ++s) {
		uint32_t x = h[s] ^ k;
		x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
		h[s] = x;
	}
}

// ---------------------------------------------------------------- synthetic data
KWG_DEV uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
KWG_DEV uint64_t synth_rnd(uint64_t seed, uint64_t stream, uint64_t ctr)
{
	return mix64(mix64(seed ^ mix64(stream)) + ctr * 0x9E3779B97F4A7C15ull);
}


// ---- 16 ASCII bases -> 32 bits of 2-bit codes (first base in the top bits) + 16 "bad" flags
// (bit j set <=> base j is not one of ACGTacgt; reference word.h:80-101).
KWG_DEV uint32_t nonzero_bytes(uint32_t v)   // 0x80 in every non-zero byte
{
	return (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}

KWG_DEV void encode4(uint32_t w, uint32_t& code8, uint32_t& bad4)
{
	const uint32_t u = w & 0xDFDFDFDFu;                      // fold lower case onto upper case
	const uint32_t x = (w >> 1) & 0x03030303u;               // A:0 C:1 G:3 T:2
	const uint32_t c = x ^ ((x >> 1) & 0x01010101u);         // A:0 C:1 G:2 T:3
	code8 = (c * 0x40100401u) >> 24;                         // byte0's code ends up in bits 7..6
	const uint32_t bad = nonzero_bytes(u ^ 0x41414141u) & nonzero_bytes(u ^ 0x43434343u) &
	                     nonzero_bytes(u ^ 0x47474747u) & nonzero_bytes(u ^ 0x54545454u);
	bad4 = (((bad >> 7) * 0x00204081u) >> 21) & 0xFu;        // byte j's flag -> bit j
}

KWG_DEV void encode16(uint4 v, uint32_t& codes, uint32_t& bad16)
{
	uint32_t c0, c1, c2, c3, b0, b1, b2, b3;
	encode4(v.x, c0, b0); encode4(v.y, c1, b1); encode4(v.z, c2, b2); encode4(v.w, c3, b3);
	codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
	bad16 = b0 | (b1 << 4) | (b2 << 8) | (b3 << 12);
}


// In-register transpose of a 32x32 bit matrix, LSB-first on both axes:
// out[b] bit i == in[i] bit b.
// Five butterfly stages.  The two coarse ones (half-words, bytes) are byte permutes, one instruction per output word;
// the fine ones exchange bit groups with two logic operations per output word (a shift and a three-input select).
KWG_DEV void transpose32(uint32_t (&a)[32])
{
#pragma unroll
	for (int k = 0; k < 16; ++k) {
		const uint32_t lo = a[k], hi = a[k + 16];
		a[k] = __byte_perm(lo, hi, 0x5410);           // low halves of both
		a[k + 16] = __byte_perm(lo, hi, 0x7632);      // high halves of both
	}
#pragma unroll
	for (int k = 0; k < 32; ++k) {
		if ((k & 8) == 0) {
			const uint32_t lo = a[k], hi = a[k + 8];
			a[k] = __byte_perm(lo, hi, 0x6240);       // bytes 0 and 2 of both
			a[k + 8] = __byte_perm(lo, hi, 0x7351);   // bytes 1 and 3 of both
		}
	}
#pragma unroll
	for (int j = 4; j >= 1; j >>= 1) {
		const uint32_t m = (j == 4) ? 0x0F0F0F0Fu : (j == 2) ? 0x33333333u : 0x55555555u;
#pragma unroll
		for (int k = 0; k < 32; ++k) {
			if ((k & j) == 0) {
				const uint32_t lo = a[k], hi = a[k | j];
				a[k] = (lo & m) | ((hi << j) & ~m);
				a[k | j] = ((lo >> j) & m) | (hi & ~m);
			}
		}
	}
}


KWG_DEV uint4 and4(uint4 a, uint4 b) { return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }

// carry-save adder on 128 columns: (s, carry) = s + a + b
KWG_DEV void csa(uint4& s, uint4& carry, const uint4 a, const uint4 b)
{
	carry.x = (s.x & a.x) | (b.x & (s.x ^ a.x)); s.x ^= a.x ^ b.x;
	carry.y = (s.y & a.y) | (b.y & (s.y ^ a.y)); s.y ^= a.y ^ b.y;
	carry.z = (s.z & a.z) | (b.z & (s.z ^ a.z)); s.z ^= a.z ^ b.z;
	carry.w = (s.w & a.w) | (b.w & (s.w ^ a.w)); s.w ^= a.w ^ b.w;
}


// ---- bit-sliced counters (search): 32 filter columns per word, plane i holds bit i of every count
// A substream counts with 4 low + 6 upper planes, i.e. up to 1023: a segment hands it at most SC_SUB_CAP k-mers
// (a multiple of the 16-k-mer Harley-Seal block), so a filter that holds EVERY k-mer of a long query cannot wrap.
constexpr uint32_t SC_SUB_CAP = 1008;      // k-mers per substream per segment (<= 2^10 - 1)
constexpr uint32_t SC_SEG_CAP = 32768;     // k-mers per segment (merged counts stay below 2^16)
KWG_DEV uint32_t search_seg_cap(uint32_t nsub)
{
	const uint32_t c = nsub * SC_SUB_CAP;
	return c < SC_SEG_CAP ? c : SC_SEG_CAP;
}
// ---- early exit of the thresholded search (search_count_kernel<NH, true>; shared with the host emulation, tests/host_emul)
// k-mers of a segment of seg_n that substream `sub` of nsub looks at (k-mer i goes to substream i % nsub)
KWG_DEV uint32_t search_sub_total(uint32_t seg_n, uint32_t sub, uint32_t nsub)
{
	return seg_n > sub ? (seg_n - sub + nsub - 1u) / nsub : 0u;
}
// ... of which it has not looked at yet after its block `blk` of 16
KWG_DEV uint32_t search_sub_left(uint32_t sub_total, uint32_t blk)
{
	const uint32_t done = 16u * (blk + 1u);
	return sub_total > done ? sub_total - done : 0u;
}
// the first block after which the bounds can fall below `need`: fewer than `need` k-mers of the segment are left
KWG_DEV uint32_t search_exit_first_blk(uint32_t seg_n, uint32_t need, uint32_t nsub)
{
	return (seg_n - (seg_n < need ? seg_n : need)) / (16u * nsub);
}
// bounds are refreshed after every second block from there on (not after the last: nothing is left to save)
KWG_DEV bool search_exit_check_at(uint32_t blk, uint32_t blk_first, uint32_t n_blk)
{
	return blk >= blk_first && ((blk - blk_first) & 1u) == 0u && blk + 1u < n_blk;
}

// total += x, where x has P planes and total has 16 (counts stay below 2^16 per segment)
template <int P>
KWG_DEV void bitsliced_add(uint32_t (&tot)[16], const uint32_t (&x)[P])
{
	uint32_t carry = 0;
#pragma unroll
	for (int i = 0; i < P; ++i) {
		const uint32_t t = tot[i];
		tot[i] = t ^ x[i] ^ carry;
		carry = (t & x[i]) | (carry & (t ^ x[i]));
	}
#pragma unroll
	for (int i = P; i < 16; ++i) {
		const uint32_t t = tot[i];
		tot[i] = t ^ carry;
		carry &= t;
	}
}

// counts of columns 4*nb .. 4*nb+3 of the word: the low byte of each count comes from planes 0..7,
// the high byte from planes 8..15; a nibble of a plane is spread to one bit per byte by a multiply.
KWG_DEV uint4 expand_counts4(const uint32_t (&tot)[16], int nb)
{
	uint32_t lo = 0, hi = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		lo += ((((tot[i] >> (4 * nb)) & 0xFu) * 0x00204081u) & 0x01010101u) << i;
		hi += ((((tot[i + 8] >> (4 * nb)) & 0xFu) * 0x00204081u) & 0x01010101u) << i;
	}
	return make_uint4((lo & 0xFFu) | ((hi & 0xFFu) << 8), ((lo >> 8) & 0xFFu) | (((hi >> 8) & 0xFFu) << 8),
		((lo >> 16) & 0xFFu) | (((hi >> 16) & 0xFFu) << 8), (lo >> 24) | ((hi >> 24) << 8));
}

// ---- k-mer windows over an encoded tile (construction): codes[] holds 16 bases per word (first
// base in the top bits), bad[]/start[] hold one flag bit per base (LSB first).
// A window [p, p+k) is a k-mer iff it has no bad base and no read starts strictly inside it.
KWG_DEV bool window_ok(const uint32_t* bad, const uint32_t* start, uint32_t p, uint32_t k)
{
	const uint32_t win_mask = (k >= 32) ? 0xFFFFFFFFu : ((1u << k) - 1u);
	const uint32_t bw = p >> 5, bo = p & 31;
	const uint32_t bad_win = __funnelshift_r(bad[bw], bad[bw + 1], bo) & win_mask;
	const uint32_t start_win = (__funnelshift_r(start[bw], start[bw + 1], bo) & win_mask) >> 1;
	return (bad_win | start_win) == 0;
}

// The same test for 32 consecutive start positions at once: bit i of the result <=> the window at position 32 v + i is a
// k-mer (1 <= k <= 32).  OR over a sliding window by doubling: R_1 = x, R_2n = R_n | R_n >> n, and any length is a sum of
// powers of two -- ~50 instructions per 32 positions instead of ~12 per position.
KWG_DEV uint64_t sliding_or64(uint64_t x, uint32_t n)        // bit p of the result = OR of bits p .. p + n - 1 of x (n <= 32: the low 32 bits are exact)
{
	uint64_t acc = 0, r = x;
	uint32_t done = 0;
#pragma unroll
	for (int i = 0; i < 6; ++i) {
		if (n & (1u << i)) { acc |= r >> done; done += 1u << i; }
		r |= r >> (1u << i);
	}
	return acc;
}

KWG_DEV uint32_t window_ok_word(const uint32_t* bad, const uint32_t* start, uint32_t v, uint32_t k)
{
	const uint64_t b = ((uint64_t)bad[v + 1] << 32) | bad[v];
	const uint64_t s = (((uint64_t)start[v + 1] << 32) | start[v]) >> 1;       // a read start strictly inside: offsets 1 .. k - 1
	return ~(uint32_t)(sliding_or64(b, k) | sliding_or64(s, k - 1));
}

KWG_DEV uint64_t window_sense(const uint32_t* codes, uint32_t p, uint32_t k)
{
	const uint32_t wi = p >> 4, off = 2 * (p & 15);
	const uint32_t w0 = codes[wi], w1 = codes[wi + 1], w2 = codes[wi + 2];
	const uint32_t hi = __funnelshift_l(w1, w0, off);
	const uint32_t lo = __funnelshift_l(w2, w1, off);
	return (((uint64_t)hi << 32) | lo) >> (64 - 2 * k);
}


// ---------------------------------------------------------------- k in 33..63 (raw mode only)
// PARITY UNPINNED: the reference stops at k = 32 (word.h:10).  The same rules on a 128-bit word (hi:lo, the 5' base most
// significant, 2k bits right-aligned); checked against oracle/kwo_raw_insert_wide, which equals the narrow restatement
// for k <= 32 and an independent pure-Python statement above (tests/test_oracle_wide.py).
struct Word128 { uint64_t hi, lo; };

KWG_DEV Word128 shr128(Word128 v, uint32_t s)            // 0 <= s < 128
{
	Word128 r;
	if (s == 0) return v;
	if (s >= 64) { r.hi = 0; r.lo = v.hi >> (s - 64); }
	else { r.hi = v.hi >> s; r.lo = (v.lo >> s) | (v.hi << (64 - s)); }
	return r;
}

KWG_DEV uint64_t reverse_groups64(uint64_t w)            // all 32 2-bit groups of a 64-bit word, order reversed
{
	const uint64_t r = __brevll(w);
	return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}

// the k bases starting at position p of the tile (codes[]: 16 bases per word, first base in the top bits), k <= 64
KWG_DEV Word128 window_sense_wide(const uint32_t* codes, uint32_t p, uint32_t k)
{
	const uint32_t wi = p >> 4, off = 2 * (p & 15);
	const uint32_t w0 = codes[wi], w1 = codes[wi + 1], w2 = codes[wi + 2], w3 = codes[wi + 3], w4 = codes[wi + 4];
	Word128 v;
	v.hi = ((uint64_t)__funnelshift_l(w1, w0, off) << 32) | __funnelshift_l(w2, w1, off);
	v.lo = ((uint64_t)__funnelshift_l(w3, w2, off) << 32) | __funnelshift_l(w4, w3, off);
	return shr128(v, 128 - 2 * k);
}

struct CanonWide {
	Word128 word;    // min(sense, antisense)
	Word128 low;     // the same k-mer with base i at bits [2i, 2i+1]
};

KWG_DEV CanonWide canonical_wide(Word128 sense, uint32_t k)
{
	// mask of the low 2k bits
	Word128 m;
	m.hi = (2 * k >= 128) ? ~0ull : ((2 * k > 64) ? ((1ull << (2 * k - 64)) - 1ull) : 0ull);
	m.lo = (2 * k >= 64) ? ~0ull : ((1ull << (2 * k)) - 1ull);
	// reverse the k groups: reverse all 64 groups of the 128-bit word, then drop the 64 - k empty ones at the bottom
	Word128 full;
	full.hi = reverse_groups64(sense.lo);
	full.lo = reverse_groups64(sense.hi);
	const Word128 rev = shr128(full, 128 - 2 * k);
	Word128 anti;
	anti.hi = ~rev.hi & m.hi; anti.lo = ~rev.lo & m.lo;
	CanonWide c;
	const bool sense_first = (sense.hi < anti.hi) || (sense.hi == anti.hi && sense.lo <= anti.lo);
	if (sense_first) { c.word = sense; c.low = rev; }
	else { c.word = anti; c.low.hi = ~sense.hi & m.hi; c.low.lo = ~sense.lo & m.lo; }      // reverse_groups(anti) == ~sense
	return c;
}

template <int NH>
KWG_DEV void murmur3_multi_wide(Word128 low, uint32_t k, uint32_t (&h)[NH])
{
	const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
#pragma unroll
	for (int s = 0; s < NH; ++s) h[s] = (uint32_t)s;
	const uint32_t nblocks = k >> 2;
	for (uint32_t i = 0; i < nblocks; ++i) {
		uint32_t k1 = ascii4((uint32_t)low.lo & 0xFFu);
		low = shr128(low, 8);
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
#pragma unroll
		for (int s = 0; s < NH; ++s) {
			uint32_t x = h[s] ^ k1;
			x = rotl32(x, 13);
			h[s] = x * 5u + 0xe6546b64u;
		}
	}
	const uint32_t rem = k & 3u;
	if (rem) {
		uint32_t k1 = ascii4((uint32_t)low.lo & 0xFFu) & ((1u << (8 * rem)) - 1u);
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
#pragma unroll
		for (int s = 0; s < NH; ++s) h[s] ^= k1;
	}
#pragma unroll
	for (int s = 0; s < NH; ++s) {
		uint32_t x = h[s] ^ k;
		x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
		h[s] = x;
	}
}

// a window [p, p+k), k <= 64, is a k-mer iff it has no bad base and no read starts strictly inside it
KWG_DEV bool window_ok_wide(const uint32_t* bad, const uint32_t* start, uint32_t p, uint32_t k)
{
	const uint64_t win_mask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
	const uint32_t bw = p >> 5, bo = p & 31;
	const uint64_t bad_win = (((uint64_t)__funnelshift_r(bad[bw + 1], bad[bw + 2], bo) << 32) | __funnelshift_r(bad[bw], bad[bw + 1], bo)) & win_mask;
	const uint64_t start_win = ((((uint64_t)__funnelshift_r(start[bw + 1], start[bw + 2], bo) << 32) | __funnelshift_r(start[bw], start[bw + 1], bo)) & win_mask) >> 1;
	return (bad_win | start_win) == 0;
}

} // namespace kwg
