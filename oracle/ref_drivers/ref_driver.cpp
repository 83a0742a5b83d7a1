// TEST INFRASTRUCTURE -- not product code.
//
// Command-line driver that calls the UNMODIFIED reference implementation (compiled in place
// from /root/reference by oracle/Makefile) so that tests, golden-vector generation and the
// bench's reference arm can run the reference's own hot path:
//   make_bloom_filter()  (reference make_bloom.cpp:76)
//   build_db()           (reference build_db.cpp:24)
//   bigsi_hash()         (reference hash.cpp:79-108)
//   ForEachDuplexWord    (reference word.h:73-104)
//   optimal_bloom_param(), approximate_max_kmers() (reference bloom.cpp:10-121)
// Nothing in here is copied from the reference; it only includes its headers.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "maestro.h"
#include "word.h"
#include "hash.h"
#include "kwage.h"

using namespace std;

// Globals the reference expects some translation unit to define (reference kwage.cpp:33-35).
int mpi_rank = 0;
int mpi_numtasks = 1;

// The reference sizes its counting filter from SRA metadata (make_bloom.cpp:109-129); the
// driver supplies that number explicitly so that both sides of a parity test see the same one.
static uint64_t g_number_of_bases = 0;
uint64_t number_of_bases(const std::string&) { return g_number_of_bases; }

// ---- shared synthetic-data generator (same arithmetic as oracle/kwage_oracle.c:kwo_rnd) ----
static inline uint64_t mix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}
static inline uint64_t kw_rnd(uint64_t seed, uint64_t stream, uint64_t ctr)
{
	return mix64(mix64(seed ^ mix64(stream)) + ctr * 0x9E3779B97F4A7C15ULL);
}

static double now_s()
{
	return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

static int usage()
{
	cerr << "ref_driver <command> ...\n"
		"  hash <k> <num_seed> <sequence>      canonical words + bigsi_hash of every valid window\n"
		"  optparam <k> <num_kmer> <p> <Lmin> <Lmax>\n"
		"  maxkmers <p> <Lmin> <Lmax>\n"
		"  make_bloom <accession> <bloom_dir> <k> <min_count> <p> <Lmin> <Lmax> <num_bp>\n"
		"  build_db <out.db> <k> <L> <num_hash> <file-with-bloom-paths>\n"
		"  gen_blooms <dir> <num_filter> <L> <k> <num_hash> <seed>   (25% dense random filters)\n"
		"  wrap_bloom <out.bloom> <accession> <k> <L> <num_hash> <raw-bits-file>\n"
		"  gen_reads_mt64 <out.reads> <seed> <num_read> <read_len>\n";
	return 2;
}

static int cmd_hash(int argc, char** argv)
{
	if(argc != 5) return usage();
	const unsigned int k = atoi(argv[2]);
	const unsigned int num_seed = atoi(argv[3]);
	const string seq = argv[4];
	vector<size_t> hv(num_seed);
	const char* b = seq.c_str();
	const char* e = b + seq.size();
	ForEachDuplexWord(b, e, k)
		if(ValidWord){
			const Word w = CanonicalWord;
			bigsi_hash(hv, w, k, MURMUR_HASH_32);
			printf("%zu %016zx %016zx %016zx", (size_t)Loc5, (size_t)SenseWord, (size_t)AntisenseWord, (size_t)w);
			for(unsigned int s = 0;s < num_seed;++s){
				// The three entry points must agree (hash.cpp:62-108)
				const size_t a = bigsi_hash(w, k, s, MURMUR_HASH_32);
				const size_t c = bigsi_hash(word_to_string(w, k), s, MURMUR_HASH_32);
				if( (a != hv[s]) || (c != hv[s]) ){
					fprintf(stderr, "hash entry points disagree\n");
					return 1;
				}
				printf(" %08zx", hv[s]);
			}
			printf(" %s\n", word_to_string(w, k).c_str());
		}
	EndWord
	return 0;
}

static int cmd_optparam(int argc, char** argv)
{
	if(argc != 7) return usage();
	try{
		const BloomParam p = optimal_bloom_param(atoi(argv[2]), strtoull(argv[3], NULL, 10), (float)atof(argv[4]),
			MURMUR_HASH_32, atoi(argv[5]), atoi(argv[6]));
		printf("%u %u\n", p.log_2_filter_len, p.num_hash);
	}
	catch(const char* err){
		printf("throw\n");
	}
	return 0;
}

static int cmd_maxkmers(int argc, char** argv)
{
	if(argc != 5) return usage();
	printf("%zu\n", approximate_max_kmers((float)atof(argv[2]), MURMUR_HASH_32, atoi(argv[3]), atoi(argv[4])));
	return 0;
}

static int cmd_make_bloom(int argc, char** argv)
{
	if(argc != 10) return usage();
	const string acc = argv[2];
	const string bloom_dir = argv[3];
	MaestroOptions opt;
	opt.kmer_len = atoi(argv[4]);
	opt.min_kmer_count = atoi(argv[5]);
	opt.false_positive_probability = (float)atof(argv[6]);
	opt.min_log_2_filter_len = atoi(argv[7]);
	opt.max_log_2_filter_len = atoi(argv[8]);
	opt.hash_func = MURMUR_HASH_32;
	opt.verbose = false;
	g_number_of_bases = strtoull(argv[9], NULL, 10);

	FilterInfo info;
	info.run_accession = str_to_accession(acc);
	BloomParam param;
	BloomProgress progress;

	const double t0 = now_s();
	const unsigned char status = make_bloom_filter(info.run_accession, info, param, progress, bloom_dir, opt);
	const double t1 = now_s();

	printf("{\"status\": %d, \"num_kmer\": %zu, \"num_bp\": %zu, \"log_2_filter_len\": %u, \"num_hash\": %u, "
		"\"log_2_counting_filter_len\": %zu, \"seconds\": %.6f, \"error\": \"%s\"}\n",
		int(status), progress.num_kmer, progress.num_bp, param.log_2_filter_len, param.num_hash,
		progress.log_2_counting_filter_len, t1 - t0, progress.error.c_str());
	return 0;
}

static int cmd_build_db(int argc, char** argv)
{
	if(argc != 7) return usage();
	BloomParam param;
	param.kmer_len = atoi(argv[3]);
	param.log_2_filter_len = atoi(argv[4]);
	param.num_hash = atoi(argv[5]);
	param.hash_func = MURMUR_HASH_32;
	deque<string> files;
	ifstream fin(argv[6]);
	string line;
	while(getline(fin, line)) if(!line.empty()) files.push_back(line);
	const double t0 = now_s();
	const bool ok = build_db(argv[2], param, files);
	const double t1 = now_s();
	printf("{\"ok\": %s, \"num_filter\": %zu, \"seconds\": %.6f}\n", ok ? "true" : "false", files.size(), t1 - t0);
	return ok ? 0 : 1;
}

static string fixture_accession(size_t i)
{
	stringstream ss;
	ss << "SRR" << (1000000 + i);
	return ss.str();
}

static int cmd_gen_blooms(int argc, char** argv)
{
	if(argc != 8) return usage();
	const string dir = argv[2];
	const size_t n = strtoull(argv[3], NULL, 10);
	BloomParam param;
	param.log_2_filter_len = atoi(argv[4]);
	param.kmer_len = atoi(argv[5]);
	param.num_hash = atoi(argv[6]);
	param.hash_func = MURMUR_HASH_32;
	const uint64_t seed = strtoull(argv[7], NULL, 10);
	for(size_t j = 0;j < n;++j){
		BloomFilter f(param);
		uint64_t* p = (uint64_t*)f.ptr();
		const size_t nw = f.num_block()/8;
		for(size_t w = 0;w < nw;++w) p[w] = kw_rnd(seed, j, 2*w) & kw_rnd(seed, j, 2*w + 1);
		f.update_crc32();
		FilterInfo info;
		info.run_accession = str_to_accession(fixture_accession(j));
		f.set_info(info);
		const string path = dir + "/" + fixture_accession(j) + ".bloom";
		ofstream fout(path.c_str(), ios::binary);
		binary_write(fout, f);
		if(!fout){ cerr << "write failed: " << path << endl; return 1; }
		printf("%s\n", path.c_str());
	}
	return 0;
}

static int cmd_wrap_bloom(int argc, char** argv)
{
	if(argc != 8) return usage();
	BloomParam param;
	param.kmer_len = atoi(argv[4]);
	param.log_2_filter_len = atoi(argv[5]);
	param.num_hash = atoi(argv[6]);
	param.hash_func = MURMUR_HASH_32;
	BloomFilter f(param);
	ifstream fin(argv[7], ios::binary);
	fin.read((char*)f.ptr(), f.num_block());
	if(!fin){ cerr << "short read of raw bits" << endl; return 1; }
	f.update_crc32();
	FilterInfo info;
	info.run_accession = str_to_accession(argv[3]);
	f.set_info(info);
	ofstream fout(argv[2], ios::binary);
	binary_write(fout, f);
	printf("%08x\n", f.get_crc32());
	return fout ? 0 : 1;
}

static int cmd_gen_reads_mt64(int argc, char** argv)
{
	// The fixture SURVEY.md section 8c pins: std::mt19937_64, base = "ACGT"[rng() & 3], row-major
	if(argc != 6) return usage();
	ofstream fout(argv[2]);
	std::mt19937_64 rng(strtoull(argv[3], NULL, 10));
	const size_t n = strtoull(argv[4], NULL, 10);
	const size_t len = strtoull(argv[5], NULL, 10);
	string s(len, 'A');
	for(size_t i = 0;i < n;++i){
		for(size_t j = 0;j < len;++j) s[j] = "ACGT"[rng() & 3];
		fout << s << '\n';
	}
	return fout ? 0 : 1;
}

int main(int argc, char** argv)
{
	if(argc < 2) return usage();
	const string cmd = argv[1];
	try{
		if(cmd == "hash") return cmd_hash(argc, argv);
		if(cmd == "optparam") return cmd_optparam(argc, argv);
		if(cmd == "maxkmers") return cmd_maxkmers(argc, argv);
		if(cmd == "make_bloom") return cmd_make_bloom(argc, argv);
		if(cmd == "build_db") return cmd_build_db(argc, argv);
		if(cmd == "gen_blooms") return cmd_gen_blooms(argc, argv);
		if(cmd == "wrap_bloom") return cmd_wrap_bloom(argc, argv);
		if(cmd == "gen_reads_mt64") return cmd_gen_reads_mt64(argc, argv);
	}
	catch(const char* err){
		cerr << "reference threw: " << err << endl;
		return 1;
	}
	catch(const std::exception& err){
		cerr << "reference threw: " << err.what() << endl;
		return 1;
	}
	return usage();
}
