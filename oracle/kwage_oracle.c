/*
 * kwage_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the algorithm of KWAGE's k-mer Bloom-filter hot path, written from
 * the behaviour of the reference (citations are into /root/reference).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can CHECK the CUDA path; the
 * product (kwage_b200/, libkwage_cuda.so) never links, loads or calls it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this file is
 * pinned against the UNMODIFIED reference compiled into oracle/_ref/ (oracle/Makefile) -- see
 * tests/test_oracle_vs_reference.py and the fixtures it generated under tests/golden/.
 *
 * Build: gcc -O2 -std=c11 -fPIC -shared kwage_oracle.c -o libkwage_oracle.so -lm -lz
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define KWO_MAX_NUM_HASH 5           /* reference bloom.h:21 */
#define KWO_MIN_NUM_HASH 1           /* reference bloom.h:20 */
#define KWO_MIN_LOG_COUNT_LEN 18     /* reference make_bloom.cpp:22 */
#define KWO_MAX_LOG_COUNT_LEN 32     /* reference make_bloom.cpp:21 */
#define KWO_COUNT_FILTER_FP 1.0e-2   /* reference make_bloom.cpp:25 */

/* ------------------------------------------------------------------------------------------ */
/* Synthetic data generators shared (same arithmetic) with kwage_b200/synth.py and the device   */
/* generator in kwage_b200/csrc/synth.cu.  Not part of the reference.                           */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

uint64_t kwo_rnd(uint64_t seed, uint64_t stream, uint64_t ctr)
{
	return mix64(mix64(seed ^ mix64(stream)) + ctr * 0x9E3779B97F4A7C15ULL);
}

/* n_reads reads of read_len bases, row-major, no separators: base p of read r takes 2 bits of
 * kwo_rnd(seed, first_read + r, p / 32). */
void kwo_gen_reads(uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len, char* out)
{
	for (uint64_t r = 0; r < n_reads; ++r) {
		uint64_t w = 0;
		for (uint32_t p = 0; p < read_len; ++p) {
			if ((p & 31) == 0) w = kwo_rnd(seed, first_read + r, p >> 5);
			out[r * read_len + p] = "ACGT"[(w >> (2 * (p & 31))) & 3];
		}
	}
}

/* 25%-dense random filter bits: 64-bit word w of filter j = rnd(seed, j, 2w) & rnd(seed, j, 2w+1) */
void kwo_gen_filter_bits(uint64_t seed, uint64_t filter, uint64_t n_words, uint64_t* out)
{
	for (uint64_t w = 0; w < n_words; ++w)
		out[w] = kwo_rnd(seed, filter, 2 * w) & kwo_rnd(seed, filter, 2 * w + 1);
}

/* std::mt19937_64 (public algorithm, Matsumoto & Nishimura) -- reproduces the fixture that
 * SURVEY.md section 8c pins: base = "ACGT"[rng() & 3], row-major. */
void kwo_gen_reads_mt64(uint64_t seed, uint64_t n_reads, uint32_t read_len, char* out)
{
	enum { NN = 312, MM = 156 };
	static const uint64_t MATRIX_A = 0xB5026F5AA96619E9ULL, UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
	uint64_t mt[NN];
	int mti;
	mt[0] = seed;
	for (mti = 1; mti < NN; ++mti) mt[mti] = 6364136223846793005ULL * (mt[mti - 1] ^ (mt[mti - 1] >> 62)) + (uint64_t)mti;
	const uint64_t total = n_reads * (uint64_t)read_len;
	for (uint64_t o = 0; o < total; ++o) {
		if (mti >= NN) {
			int i;
			uint64_t x;
			for (i = 0; i < NN - MM; ++i) {
				x = (mt[i] & UM) | (mt[i + 1] & LM);
				mt[i] = mt[i + MM] ^ (x >> 1) ^ ((x & 1ULL) ? MATRIX_A : 0ULL);
			}
			for (; i < NN - 1; ++i) {
				x = (mt[i] & UM) | (mt[i + 1] & LM);
				mt[i] = mt[i + (MM - NN)] ^ (x >> 1) ^ ((x & 1ULL) ? MATRIX_A : 0ULL);
			}
			x = (mt[NN - 1] & UM) | (mt[0] & LM);
			mt[NN - 1] = mt[MM - 1] ^ (x >> 1) ^ ((x & 1ULL) ? MATRIX_A : 0ULL);
			mti = 0;
		}
		uint64_t x = mt[mti++];
		x ^= (x >> 29) & 0x5555555555555555ULL;
		x ^= (x << 17) & 0x71D67FFFEDA60000ULL;
		x ^= (x << 37) & 0xFFF7EEE000000000ULL;
		x ^= (x >> 43);
		out[o] = "ACGT"[x & 3];
	}
}

/* ------------------------------------------------------------------------------------------ */
/* k-mer digestion: reference word.h:73-104 (ForEachDuplexWord), 161-172 (ValidWord, SenseWord,  */
/* AntisenseWord, CanonicalWord, Loc5) and word.cpp:9-23 (kmer_word_mask).                       */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t kmer_word_mask(uint32_t k)
{
	return (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
}

typedef struct {
	uint64_t sense, anti, mask;
	uint32_t word_len, k;
} kwo_roll_t;

static inline void roll_init(kwo_roll_t* s, uint32_t k)
{
	s->sense = s->anti = 0;
	s->mask = kmer_word_mask(k);
	s->word_len = 0;
	s->k = k;
}

/* Shift one base in; returns 1 when the current window is a valid k-mer. */
static inline int roll_push(kwo_roll_t* s, char c)
{
	const uint32_t comp_shift = 2 * (s->k - 1);
	uint64_t code;
	++s->word_len;
	switch (c) {
		case 'A': case 'a': code = 0; break;
		case 'C': case 'c': code = 1; break;
		case 'G': case 'g': code = 2; break;
		case 'T': case 't': code = 3; break;
		default:
			/* word.h:98-100: the words are left untouched and the run length restarts */
			s->word_len = 0;
			return 0;
	}
	s->sense = (s->sense << 2) | code;
	s->anti = (s->anti >> 2) | ((3 - code) << comp_shift);
	return s->word_len >= s->k;
}

static inline uint64_t roll_canonical(const kwo_roll_t* s)
{
	const uint64_t a = s->sense & s->mask, b = s->anti & s->mask;
	return a < b ? a : b;
}

/* Canonical words of every valid window of seq[0,len), in stream order.  out_words / out_loc5 may
 * be NULL (count only).  Returns the number of valid windows. */
uint64_t kwo_canonical_kmers(const char* seq, uint64_t len, uint32_t k, uint64_t* out_words, uint64_t* out_loc5)
{
	kwo_roll_t s;
	uint64_t n = 0;
	roll_init(&s, k);
	for (uint64_t i = 0; i < len; ++i) {
		if (roll_push(&s, seq[i])) {
			if (out_words) out_words[n] = roll_canonical(&s);
			if (out_loc5) out_loc5[n] = (i + 1) - k;
			++n;
		}
	}
	return n;
}

/* ------------------------------------------------------------------------------------------ */
/* hash: reference hash.cpp:176-234 (murmur_hash32 over the ASCII of a 2-bit word; identical to  */
/* the AVX2 multi-seed version hash.cpp:239-332 and the string version hash.cpp:114-170).        */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline uint32_t base_ascii(uint64_t w, uint32_t k, uint32_t index)
{
	return (uint32_t)"ACGT"[(w >> (2 * (k - 1 - index))) & 3];
}

uint32_t kwo_murmur3_word(uint64_t w, uint32_t k, uint32_t seed)
{
	const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
	const uint32_t nblocks = k / 4;
	uint32_t h1 = seed, off = 0;
	for (uint32_t i = 0; i < nblocks; ++i, off += 4) {
		uint32_t k1 = (base_ascii(w, k, off + 3) << 24) | (base_ascii(w, k, off + 2) << 16) |
		              (base_ascii(w, k, off + 1) << 8) | base_ascii(w, k, off);
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
		h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
	}
	uint32_t k1 = 0;
	switch (k & 3) {
		case 3: k1 ^= base_ascii(w, k, off + 2) << 16; /* fall through */
		case 2: k1 ^= base_ascii(w, k, off + 1) << 8;  /* fall through */
		case 1: k1 ^= base_ascii(w, k, off);
			k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2; h1 ^= k1;
	}
	h1 ^= k;
	h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
	return h1;
}

/* Textbook MurmurHash3_x86_32 over bytes, used by the tests to confirm the statement
 * "hash == murmur3 of the ASCII k-mer" (SURVEY.md F6). */
uint32_t kwo_murmur3_bytes(const uint8_t* data, uint32_t len, uint32_t seed)
{
	const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
	uint32_t h1 = seed, i;
	for (i = 0; i + 4 <= len; i += 4) {
		uint32_t k1 = (uint32_t)data[i] | ((uint32_t)data[i + 1] << 8) | ((uint32_t)data[i + 2] << 16) | ((uint32_t)data[i + 3] << 24);
		k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2;
		h1 ^= k1; h1 = rotl32(h1, 13); h1 = h1 * 5 + 0xe6546b64u;
	}
	uint32_t k1 = 0;
	switch (len & 3) {
		case 3: k1 ^= (uint32_t)data[i + 2] << 16; /* fall through */
		case 2: k1 ^= (uint32_t)data[i + 1] << 8;  /* fall through */
		case 1: k1 ^= data[i];
			k1 *= c1; k1 = rotl32(k1, 15); k1 *= c2; h1 ^= k1;
	}
	h1 ^= len;
	h1 ^= h1 >> 16; h1 *= 0x85ebca6bu; h1 ^= h1 >> 13; h1 *= 0xc2b2ae35u; h1 ^= h1 >> 16;
	return h1;
}

/* ------------------------------------------------------------------------------------------ */
/* bit vectors: reference bloom.h:131-163 -- bit i lives in byte i/8 at bit position i%8.        */
/* ------------------------------------------------------------------------------------------ */
static inline void set_bit(uint8_t* buf, uint64_t i) { buf[i >> 3] |= (uint8_t)(1u << (i & 7)); }
static inline int get_bit(const uint8_t* buf, uint64_t i) { return (buf[i >> 3] >> (i & 7)) & 1; }

uint32_t kwo_crc32(const uint8_t* buf, uint64_t len)
{
	/* reference bloom.cpp:328-336: zlib CRC-32 seeded by crc32_z(0, NULL, 0) */
	uLong c = crc32_z(0L, Z_NULL, 0);
	return (uint32_t)crc32_z(c, buf, (z_size_t)len);
}

/* ------------------------------------------------------------------------------------------ */
/* Raw insert with fixed parameters: the ground-truth construction of the reference's own test   */
/* rig, bloom_test.cpp:268-275: set_bit(bigsi_hash(kmer, k, h) % filter_len) for h < num_hash.    */
/* bits must hold 2^log2_len / 8 bytes (not cleared here, so calls accumulate).                  */
/* ------------------------------------------------------------------------------------------ */
uint64_t kwo_raw_insert(const char* bases, const uint64_t* offsets, uint64_t n_reads, uint32_t k,
	uint32_t num_hash, uint32_t log2_len, uint8_t* bits)
{
	const uint64_t mask = (1ULL << log2_len) - 1ULL;
	uint64_t n = 0;
	for (uint64_t r = 0; r < n_reads; ++r) {
		kwo_roll_t s;
		roll_init(&s, k);
		for (uint64_t i = offsets[r]; i < offsets[r + 1]; ++i) {
			if (roll_push(&s, bases[i])) {
				const uint64_t w = roll_canonical(&s);
				for (uint32_t h = 0; h < num_hash; ++h) set_bit(bits, kwo_murmur3_word(w, k, h) & mask);
				++n;
			}
		}
	}
	return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Bloom parameter search: reference bloom.cpp:10-68 and 72-121.  float/double mixing kept.      */
/* ------------------------------------------------------------------------------------------ */
/* returns 0 and fills log2_len / num_hash, or -1 where the reference throws */
/* ---------------------------------------------------------------------------------------------
 * k in 33..63 -- PARITY UNPINNED.  The reference stops at k = 32 (word.h:10: a k-mer is one 64-bit Word), so there is no
 * reference output to check this against; BASELINE.json configs[4] asks for k up to 63 all the same.  What follows is the
 * reference's rule set carried over to a 128-bit word, nothing else changed: 2 bits per base with the 5' base most
 * significant (word.h:73-104), any byte other than ACGTacgt ends the current run of valid bases (word.h:98-100), canonical
 * = min(sense, reverse complement) as an unsigned integer (word.h:163-165), murmur3_x86_32 over the k ASCII bytes of the
 * canonical k-mer in 5'->3' order, seeds 0..n-1 (hash.cpp:176-234), bit (hash & (2^L - 1)) set LSB first.
 * For k <= 32 it gives exactly what kwo_raw_insert gives (tests/test_oracle_wide.py), which is as far as pinning can go.
 * ------------------------------------------------------------------------------------------- */
typedef unsigned __int128 kwo_u128;

static uint32_t murmur3_wide(kwo_u128 w, uint32_t k, uint32_t seed)
{
	uint8_t ascii[64];
	for (uint32_t i = 0; i < k; ++i) ascii[i] = (uint8_t)"ACGT"[(unsigned)(w >> (2 * (k - 1 - i))) & 3u];
	return kwo_murmur3_bytes(ascii, k, seed);
}

uint64_t kwo_raw_insert_wide(const char* bases, const uint64_t* offsets, uint64_t n_reads, uint32_t k,
	uint32_t num_hash, uint32_t log2_len, uint8_t* bits)
{
	const uint64_t mask = (log2_len >= 64) ? ~0ULL : ((1ULL << log2_len) - 1ULL);
	const kwo_u128 kmask = (k >= 64) ? ~(kwo_u128)0 : ((((kwo_u128)1) << (2 * k)) - 1);
	uint64_t n = 0;
	for (uint64_t r = 0; r < n_reads; ++r) {
		kwo_u128 sense = 0, anti = 0;
		uint32_t valid = 0;
		for (uint64_t i = offsets[r]; i < offsets[r + 1]; ++i) {
			unsigned code;
			switch (bases[i]) {
				case 'A': case 'a': code = 0; break;
				case 'C': case 'c': code = 1; break;
				case 'G': case 'g': code = 2; break;
				case 'T': case 't': code = 3; break;
				default: valid = 0; sense = 0; anti = 0; continue;
			}
			sense = ((sense << 2) | code) & kmask;
			anti = (anti >> 2) | ((kwo_u128)(3u - code) << (2 * (k - 1)));
			if (++valid >= k) {
				const kwo_u128 w = sense < anti ? sense : anti;
				for (uint32_t h = 0; h < num_hash; ++h) set_bit(bits, murmur3_wide(w, k, h) & mask);
				++n;
			}
		}
	}
	return n;
}

int kwo_optimal_bloom_param(uint64_t num_kmer, float p_max, uint32_t min_log2, uint32_t max_log2,
	uint32_t* log2_len, uint32_t* num_hash)
{
	if (num_kmer == 0) return -1;
	int valid = 0;
	for (uint32_t L = min_log2; L <= max_log2; ++L) {
		float best_p = 10.0f;
		for (uint32_t h = KWO_MIN_NUM_HASH; h <= KWO_MAX_NUM_HASH; ++h) {
			const uint64_t len = 1ULL << L;
			/* bloom.cpp:47 -- note m_num_kmer*num_hash is an integer product converted to double */
			const double p = pow(1.0 - pow(1.0 - 1.0 / len, (double)(num_kmer * h)), (double)h);
			if ((p <= p_max) && (p < best_p)) {
				best_p = (float)p;
				*num_hash = h;
				valid = 1;
			}
		}
		if (valid) {
			*log2_len = L;
			return 0;
		}
	}
	return -1;
}

uint64_t kwo_approximate_max_kmers(float p_max, uint32_t min_log2, uint32_t max_log2)
{
	for (uint32_t lk = 1; lk < 64; ++lk) {
		const uint64_t num_kmer = 1ULL << lk;
		int valid = 0;
		for (uint32_t L = min_log2; (L <= max_log2) && !valid; ++L) {
			const float best_p = 10.0f;
			for (uint32_t h = KWO_MIN_NUM_HASH; (h <= KWO_MAX_NUM_HASH) && !valid; ++h) {
				const uint64_t len = 1ULL << L;
				const double p = pow(1.0 - pow(1.0 - 1.0 / len, (double)(num_kmer * h)), (double)h);
				if ((p <= p_max) && (p < best_p)) valid = 1;
			}
		}
		if (!valid) return num_kmer;
	}
	return 0xFFFFFFFFFFFFFFFFULL;
}

/* Counting-filter length from the metadata base count: reference make_bloom.cpp:104-129 */
uint32_t kwo_counting_log2_len(uint64_t num_bp)
{
	uint64_t L = KWO_MAX_LOG_COUNT_LEN;
	if (num_bp > 0) {
		const double counting_length = 1.0 / (1.0 - pow(1.0 - pow(KWO_COUNT_FILTER_FP, 1.0 / 4.0), 1.0 / (2 * num_bp)));
		L = (uint64_t)ceil(log(counting_length) / log(2.0));
		if (L > KWO_MAX_LOG_COUNT_LEN) L = KWO_MAX_LOG_COUNT_LEN;
		if (L < KWO_MIN_LOG_COUNT_LEN) L = KWO_MIN_LOG_COUNT_LEN;
	}
	return (uint32_t)L;
}

/* ------------------------------------------------------------------------------------------ */
/* Production construction: reference make_bloom.cpp:63-69 (CountingBloom: low nibble "first",   */
/* high nibble "second" with GCC bit-field order), 151-166 (tables), 506-621 (count_words) and   */
/* 337-354 (fold).  Sequential and order-dependent, exactly like the reference, for every        */
/* min_kmer_count in [1,15].                                                                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t k, min_count, log2_count_len, log2_max_len;
	uint64_t num_valid;
	uint8_t* count;                     /* 2^log2_count_len bytes, two nibbles each */
	uint8_t* valid[KWO_MAX_NUM_HASH];   /* 5 bit vectors of 2^log2_max_len bits */
} kwo_builder_t;

void kwo_builder_destroy(kwo_builder_t* b)
{
	if (!b) return;
	free(b->count);
	for (int h = 0; h < KWO_MAX_NUM_HASH; ++h) free(b->valid[h]);
	free(b);
}

kwo_builder_t* kwo_builder_create(uint32_t k, uint32_t min_count, uint32_t log2_count_len, uint32_t log2_max_len)
{
	kwo_builder_t* b = (kwo_builder_t*)calloc(1, sizeof(kwo_builder_t));
	if (!b) return NULL;
	b->k = k; b->min_count = min_count; b->log2_count_len = log2_count_len; b->log2_max_len = log2_max_len;
	b->count = (uint8_t*)calloc(1ULL << log2_count_len, 1);
	int ok = b->count != NULL;
	const uint64_t vbytes = (log2_max_len >= 3) ? (1ULL << (log2_max_len - 3)) : 1;
	for (int h = 0; h < KWO_MAX_NUM_HASH; ++h) {
		b->valid[h] = (uint8_t*)calloc(vbytes, 1);
		ok = ok && (b->valid[h] != NULL);
	}
	if (!ok) { kwo_builder_destroy(b); return NULL; }
	return b;
}

static inline uint32_t nib_first(const uint8_t* c, uint64_t i) { return c[i] & 15u; }
static inline uint32_t nib_second(const uint8_t* c, uint64_t i) { return c[i] >> 4; }
/* 4-bit bit-field increment wraps 15 -> 0 without touching the other nibble */
static inline void inc_first(uint8_t* c, uint64_t i) { c[i] = (uint8_t)((c[i] & 0xF0u) | ((c[i] + 1u) & 0x0Fu)); }
static inline void inc_second(uint8_t* c, uint64_t i) { c[i] = (uint8_t)((c[i] & 0x0Fu) | ((c[i] + 0x10u) & 0xF0u)); }

/* count_words() over one fragment: make_bloom.cpp:506-621 */
static void builder_add_fragment(kwo_builder_t* b, const char* seq, uint64_t len)
{
	const uint64_t cmask = (1ULL << b->log2_count_len) - 1ULL;
	const uint64_t smask = (1ULL << b->log2_max_len) - 1ULL;
	kwo_roll_t s;
	roll_init(&s, b->k);
	for (uint64_t i = 0; i < len; ++i) {
		if (!roll_push(&s, seq[i])) continue;
		const uint64_t w = roll_canonical(&s);
		uint64_t hv[KWO_MAX_NUM_HASH];
		for (uint32_t h = 0; h < KWO_MAX_NUM_HASH; ++h) hv[h] = kwo_murmur3_word(w, b->k, h);
		const uint32_t f0 = nib_first(b->count, hv[0] & cmask), f1 = nib_first(b->count, hv[1] & cmask);
		const uint32_t s0 = nib_second(b->count, hv[2] & cmask), s1 = nib_second(b->count, hv[3] & cmask);
		uint32_t mn = f0;
		if (f1 < mn) mn = f1;
		if (s0 < mn) mn = s0;
		if (s1 < mn) mn = s1;
		if (mn < b->min_count) {
			if (mn == b->min_count - 1) {
				++b->num_valid;
				for (uint32_t h = 0; h < KWO_MAX_NUM_HASH; ++h) set_bit(b->valid[h], hv[h] & smask);
			}
			/* conservative update: the comparisons use the values read BEFORE any increment, so
			 * two hashes that land on the same slot increment it twice (make_bloom.cpp:586-601) */
			if (f0 == mn) inc_first(b->count, hv[0] & cmask);
			if (f1 == mn) inc_first(b->count, hv[1] & cmask);
			if (s0 == mn) inc_second(b->count, hv[2] & cmask);
			if (s1 == mn) inc_second(b->count, hv[3] & cmask);
		}
	}
}

void kwo_builder_add_reads(kwo_builder_t* b, const char* bases, const uint64_t* offsets, uint64_t n_reads)
{
	for (uint64_t r = 0; r < n_reads; ++r) builder_add_fragment(b, bases + offsets[r], offsets[r + 1] - offsets[r]);
}

/* Same with the reference's early-out: after every fragment make_bloom_filter() returns
 * STATUS_BLOOM_INVALID as soon as num_kmer exceeds max_num_kmer (make_bloom.cpp:208-214,246-252,
 * 288-294).  Returns the number of reads consumed (== n_reads when the limit was never crossed). */
uint64_t kwo_builder_add_reads_limit(kwo_builder_t* b, const char* bases, const uint64_t* offsets, uint64_t n_reads,
	uint64_t max_num_kmer)
{
	for (uint64_t r = 0; r < n_reads; ++r) {
		builder_add_fragment(b, bases + offsets[r], offsets[r + 1] - offsets[r]);
		if (max_num_kmer < b->num_valid) return r + 1;
	}
	return n_reads;
}

uint64_t kwo_builder_num_valid(const kwo_builder_t* b) { return b->num_valid; }

/* fold: make_bloom.cpp:337-354.  out_bits holds 2^log2_len/8 bytes and is overwritten. */
void kwo_builder_finalize(const kwo_builder_t* b, uint32_t log2_len, uint32_t num_hash, uint8_t* out_bits)
{
	const uint64_t nsrc = (b->log2_max_len >= 3) ? (1ULL << (b->log2_max_len - 3)) : 1;
	const uint64_t ndst = (log2_len >= 3) ? (1ULL << (log2_len - 3)) : 1;
	memset(out_bits, 0, ndst);
	for (uint32_t h = 0; h < num_hash; ++h)
		for (uint64_t i = 0; i < nsrc; i += ndst)
			for (uint64_t j = 0; j < ndst; ++j) out_bits[j] |= b->valid[h][i + j];
}

/* ------------------------------------------------------------------------------------------ */
/* Transposition: reference build_db.cpp:259-304.  dest (chunk_bits * ceil(n/8) bytes) is fully   */
/* overwritten; slice k holds bit j of filter j at byte j/8, bit j%8; padding bits are zero.      */
/* ------------------------------------------------------------------------------------------ */
void kwo_transpose(const uint8_t* const* filter_chunks, uint32_t n_filters, uint64_t chunk_bits, uint8_t* dest)
{
	const uint64_t bytes_per_slice = n_filters / 8 + ((n_filters % 8) ? 1 : 0);
	memset(dest, 0, chunk_bits * bytes_per_slice);
	for (uint32_t j = 0; j < n_filters; ++j) {
		const uint8_t* src = filter_chunks[j];
		for (uint64_t kbit = 0; kbit < chunk_bits; ++kbit)
			if (get_bit(src, kbit)) dest[bytes_per_slice * kbit + (j >> 3)] |= (uint8_t)(1u << (j & 7));
	}
}

/* ------------------------------------------------------------------------------------------ */
/* Search: reference kwage.cpp:340-541 over an in-memory slice region (the reference seeks in the */
/* file; slices = the bytes that follow the 44-byte header).                                      */
/* ------------------------------------------------------------------------------------------ */
static int cmp_u64(const void* a, const void* b)
{
	const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	return (x > y) - (x < y);
}

/* Sorted unique canonical k-mers of a query (kwage.cpp:352-366).  out must hold len words. */
uint64_t kwo_query_kmers(const char* query, uint64_t len, uint32_t k, uint64_t* out)
{
	uint64_t n = kwo_canonical_kmers(query, len, k, out, NULL);
	if (n == 0) return 0;
	qsort(out, n, sizeof(uint64_t), cmp_u64);
	uint64_t u = 1;
	for (uint64_t i = 1; i < n; ++i)
		if (out[i] != out[u - 1]) out[u++] = out[i];
	return u;
}

/* Full per-filter counts of query k-mers whose num_hash slices all have the filter's bit set
 * (kwage.cpp:404-433 without the early exit).  counts[n_filters] is overwritten.
 * Returns the number of unique query k-mers. */
uint64_t kwo_search_counts(const uint8_t* slices, uint32_t n_filters, uint32_t log2_len, uint32_t num_hash,
	uint32_t k, const char* query, uint64_t len, uint32_t* counts)
{
	const uint64_t slice_size = n_filters / 8 + ((n_filters % 8) ? 1 : 0);
	const uint64_t filter_len = 1ULL << log2_len;
	uint64_t* kmers = (uint64_t*)malloc((len ? len : 1) * sizeof(uint64_t));
	uint8_t* match = (uint8_t*)malloc(slice_size ? slice_size : 1);
	const uint64_t n = kwo_query_kmers(query, len, k, kmers);
	memset(counts, 0, (size_t)n_filters * sizeof(uint32_t));
	for (uint64_t i = 0; i < n; ++i) {
		memset(match, 0xFF, slice_size);
		for (uint32_t h = 0; h < num_hash; ++h) {
			const uint8_t* s = slices + (kwo_murmur3_word(kmers[i], k, h) % filter_len) * slice_size;
			for (uint64_t b = 0; b < slice_size; ++b) match[b] &= s[b];
		}
		for (uint32_t f = 0; f < n_filters; ++f) counts[f] += (uint32_t)get_bit(match, f);
	}
	free(kmers);
	free(match);
	return n;
}

/* The reference's search() including its float threshold arithmetic and both early exits
 * (kwage.cpp:349,373-389,397,404-483,489-503).  Writes matching filter ids and their reported
 * num_match; returns the number of matches.  *n_query_kmers gets the unique k-mer count. */
uint64_t kwo_search_matches(const uint8_t* slices, uint32_t n_filters, uint32_t log2_len, uint32_t num_hash,
	uint32_t k, const char* query, uint64_t len, float threshold,
	uint32_t* hit_filter, uint32_t* hit_num_match, uint32_t* n_query_kmers)
{
	const uint64_t slice_size = n_filters / 8 + ((n_filters % 8) ? 1 : 0);
	const uint64_t filter_len = 1ULL << log2_len;
	const int complete = (threshold == 1.0f);
	uint64_t* kmers = (uint64_t*)malloc((len ? len : 1) * sizeof(uint64_t));
	const uint64_t n64 = kwo_query_kmers(query, len, k, kmers);
	const unsigned int n = (unsigned int)n64;
	*n_query_kmers = n;
	if (n == 0) { free(kmers); return 0; }

	uint8_t* mask = (uint8_t*)malloc(slice_size);
	uint8_t* match = (uint8_t*)malloc(slice_size);
	uint32_t* count = (uint32_t*)calloc(n_filters, sizeof(uint32_t));
	unsigned int query_threshold = 0;
	if (complete) memset(mask, 0xFF, slice_size);
	else query_threshold = (unsigned int)(threshold * n);            /* kwage.cpp:388 (float product) */
	const uint64_t mid = (uint64_t)((1.0f - threshold) * n);          /* kwage.cpp:397 */

	for (uint64_t i = 0; i < n; ++i) {
		memset(match, 0xFF, slice_size);
		for (uint32_t h = 0; h < num_hash; ++h) {
			const uint8_t* s = slices + (kwo_murmur3_word(kmers[i], k, h) % filter_len) * slice_size;
			for (uint64_t b = 0; b < slice_size; ++b) match[b] &= s[b];
		}
		if (complete) {
			for (uint64_t b = 0; b < slice_size; ++b) mask[b] &= match[b];
		} else {
			for (uint32_t f = 0; f < n_filters; ++f) count[f] += (uint32_t)get_bit(match, f);
		}
		if (i >= mid) {
			if (complete) {
				/* max_bit(): any valid bit still set?  bloom.h:333-360 */
				int any = 0;
				for (uint32_t f = 0; f < n_filters && !any; ++f) any = get_bit(mask, f);
				if (!any) break;
			} else {
				uint32_t mx = 0;
				for (uint32_t f = 0; f < n_filters; ++f) if (count[f] > mx) mx = count[f];
				if ((uint64_t)mx + (n - i) < query_threshold) break;   /* kwage.cpp:478-481 */
			}
		}
	}

	uint64_t nh = 0;
	for (uint32_t f = 0; f < n_filters; ++f) {
		const int hit = complete ? get_bit(mask, f) : (count[f] >= query_threshold);
		if (hit) {
			hit_filter[nh] = f;
			hit_num_match[nh] = complete ? n : count[f];
			++nh;
		}
	}
	free(kmers); free(mask); free(match); free(count);
	return nh;
}
