// kwage_tools -- make_bloom / build_db front ends of the host layer (the roles of the reference's
// bff.cpp test rig and of a maestro worker's two work items, without MPI):
//   kwage_tools make_bloom [options] <accession | reads file>...     -> <bloom dir>/<ACC>.bloom
//   kwage_tools build_db -o <out.db> <a.bloom> <b.bloom> ...          (parameters come from the first file)
// Options of make_bloom mirror maestro's (options.cpp:404-820): -k, --min-kmer-count, -p,
// --len.min, --len.max, --bloom <dir>; plus --device and --num-bases (metadata override).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <map>
#include <set>

#include "kwage_host.h"

using namespace kwage;

static uint64_t g_num_bases = 0;
static uint64_t fixed_num_bases(const std::string&) { return g_num_bases; }

static std::string accession_from_path(const std::string& p, size_t index)
{
	std::string base = p.substr(p.find_last_of('/') == std::string::npos ? 0 : p.find_last_of('/') + 1);
	base = base.substr(0, base.find('.'));
	try { str_to_accession(base); return base; }
	catch (...) { char buf[32]; std::snprintf(buf, sizeof(buf), "SRR%07zu", 9000000 + index); return buf; }
}

static int cmd_make_bloom(int argc, char** argv)
{
	MaestroOptions opt;
	std::string bloom_dir = ".";
	std::deque<std::string> inputs;
	bool have_num_bases = false;
	for (int i = 2; i < argc; ++i) {
		const std::string a = argv[i];
		if (a == "-k" && i + 1 < argc) opt.kmer_len = std::atoi(argv[++i]);
		else if (a == "--min-kmer-count" && i + 1 < argc) opt.min_kmer_count = std::atoi(argv[++i]);
		else if (a == "-p" && i + 1 < argc) opt.false_positive_probability = (float)std::atof(argv[++i]);
		else if (a == "--len.min" && i + 1 < argc) opt.min_log_2_filter_len = std::atoi(argv[++i]);
		else if (a == "--len.max" && i + 1 < argc) opt.max_log_2_filter_len = std::atoi(argv[++i]);
		else if (a == "--bloom" && i + 1 < argc) bloom_dir = argv[++i];
		else if (a == "--device" && i + 1 < argc) opt.device = std::atoi(argv[++i]);
		else if (a == "--num-bases" && i + 1 < argc) { g_num_bases = std::strtoull(argv[++i], NULL, 10); have_num_bases = true; }
		else inputs.push_back(a);
	}
	// argument limits of the reference (options.cpp:716-782)
	if (opt.kmer_len < 1 || opt.kmer_len > 32) { std::cerr << "Please specify 1 <= kmer length <= 32" << std::endl; return EXIT_FAILURE; }
	if (opt.min_kmer_count > KWAGE_MAX_COUNT) { std::cerr << "Please specify a min kmer count <= 15" << std::endl; return EXIT_FAILURE; }
	if (!(opt.false_positive_probability > 0.0f && opt.false_positive_probability < 1.0f)) { std::cerr << "Please specify 0 < p < 1" << std::endl; return EXIT_FAILURE; }
	if (opt.min_log_2_filter_len > opt.max_log_2_filter_len || opt.max_log_2_filter_len > 32) { std::cerr << "Please specify len.min <= len.max <= 32" << std::endl; return EXIT_FAILURE; }
	if (have_num_bases) set_number_of_bases_hook(fixed_num_bases);
	int failures = 0;
	for (size_t i = 0; i < inputs.size(); ++i) {
		const std::string acc = accession_from_path(inputs[i], i);
		FilterInfo info;
		info.run_accession = str_to_accession(acc);
		BloomParam param;
		BloomProgress progress;
		const auto t0 = std::chrono::steady_clock::now();
		unsigned char status;
		try {
			ReadSource* src = open_read_collection(inputs[i]);
			status = make_bloom_filter(*src, number_of_bases(acc), info.run_accession, info, param, progress, bloom_dir, opt);
			delete src;
		}
		catch (const char* error) { progress.error = error; status = STATUS_BLOOM_FAIL; }
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		std::cout << "{\"accession\": \"" << acc << "\", \"status\": " << int(status) << ", \"num_kmer\": " << progress.num_kmer
		          << ", \"num_bp\": " << progress.num_bp << ", \"log_2_filter_len\": " << param.log_2_filter_len << ", \"num_hash\": "
		          << param.num_hash << ", \"log_2_counting_filter_len\": " << progress.log_2_counting_filter_len << ", \"seconds\": " << sec
		          << ", \"error\": \"" << progress.error << "\"}" << std::endl;
		failures += (status == STATUS_BLOOM_FAIL);
	}
	return failures ? EXIT_FAILURE : EXIT_SUCCESS;
}

static int cmd_build_db(int argc, char** argv)
{
	std::string out;
	std::deque<std::string> files;
	for (int i = 2; i < argc; ++i) {
		const std::string a = argv[i];
		if (a == "-o" && i + 1 < argc) out = argv[++i];
		else if (a == "--device" && i + 1 < argc) set_build_db_device(std::atoi(argv[++i]));
		else if (a == "--list" && i + 1 < argc) {
			std::ifstream fin(argv[++i]);
			std::string line;
			while (std::getline(fin, line)) if (!line.empty()) files.push_back(line);
		}
		else files.push_back(a);
	}
	if (out.empty() || files.empty()) { std::cerr << "build_db -o <out.db> [--list <file>] <bloom files...>" << std::endl; return EXIT_FAILURE; }
	BloomFileHeader head;
	try {
		std::ifstream fin(files[0].c_str(), std::ios::binary);
		if (!fin) throw __FILE__ ":build_db: Unable to open Bloom filter file";
		read_bloom_header(fin, head);
	}
	catch (const char* error) { std::cerr << "Caught the error " << error << std::endl; return EXIT_FAILURE; }
	const auto t0 = std::chrono::steady_clock::now();
	const bool ok = build_db(out, head.param, files);
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	std::cout << "{\"ok\": " << (ok ? "true" : "false") << ", \"num_filter\": " << files.size() << ", \"seconds\": " << sec << "}" << std::endl;
	return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}

// merge_db <a.db> <b.db> ...: the reference's merge_db main (merge_db.cpp:26-277).  Files that are not full are grouped by
// Bloom parameters; within a group the two with the fewest filters are merged (the smaller into the larger) until at
// most one partially filled file is left.
static int cmd_merge_db(int argc, char** argv)
{
	int device = 0;
	std::deque<std::string> inputs;
	for (int i = 2; i < argc; ++i) {
		const std::string a = argv[i];
		if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
		else inputs.push_back(a);
	}
	if (inputs.size() < 2) { std::cerr << "Please specify 2 or more database files to merge" << std::endl; return EXIT_SUCCESS; }
	try {
		struct Key {
			uint32_t k, L, h; int func;
			bool operator<(const Key& r) const { return std::make_pair(std::make_pair(k, L), std::make_pair(h, func)) < std::make_pair(std::make_pair(r.k, r.L), std::make_pair(r.h, r.func)); }
		};
		std::map<Key, std::deque<std::pair<size_t, std::string> > > groups;
		std::set<std::string> seen;
		for (size_t i = 0; i < inputs.size(); ++i) {
			std::ifstream fin(inputs[i].c_str(), std::ios::binary);
			if (!fin) { std::cerr << "Unable to open " << inputs[i] << " for reading" << std::endl; return EXIT_FAILURE; }
			DBFileHeader h;
			binary_read(fin, h);
			if (!fin) { std::cerr << "Unable to read database header" << std::endl; return EXIT_FAILURE; }
			if (max_filters_per_database_file(h.log_2_filter_len) <= h.num_filter) continue;      // full: not a merge candidate
			if (!seen.insert(inputs[i]).second) { std::cerr << inputs[i] << " appears more than once in the input file list" << std::endl; return EXIT_FAILURE; }
			const Key key = {h.kmer_len, h.log_2_filter_len, h.num_hash, h.hash_func};
			groups[key].push_back(std::make_pair(size_t(h.num_filter), inputs[i]));
		}
		std::cerr << "Found " << groups.size() << " distinct Bloom parameter groups" << std::endl;
		for (std::map<Key, std::deque<std::pair<size_t, std::string> > >::iterator g = groups.begin(); g != groups.end(); ++g) {
			std::deque<std::pair<size_t, std::string> >& files = g->second;
			std::sort(files.begin(), files.end());
			const size_t max_f = max_filters_per_database_file(g->first.L);
			while (files.size() > 1) {
				const std::string file_small = files.front().second;
				files.pop_front();
				const std::string file_large = files.front().second;
				files.pop_front();
				std::cerr << "\tmerging:\n\t\t" << file_small << "\n\t\t" << file_large << std::endl;
				const std::pair<size_t, std::string> rest = merge_database_files(file_large, file_small, max_f, device);
				if (rest.first > 0) { files.push_back(rest); std::sort(files.begin(), files.end()); }
			}
		}
	}
	catch (const char* error) { std::cerr << "Caught the error " << error << std::endl; return EXIT_FAILURE; }
	catch (...) { std::cerr << "Caught an unhandled error" << std::endl; return EXIT_FAILURE; }
	return EXIT_SUCCESS;
}

int main(int argc, char** argv)
{
	if (argc >= 2 && std::strcmp(argv[1], "merge_db") == 0) return cmd_merge_db(argc, argv);
	if (argc >= 2 && std::strcmp(argv[1], "make_bloom") == 0) return cmd_make_bloom(argc, argv);
	if (argc >= 2 && std::strcmp(argv[1], "build_db") == 0) return cmd_build_db(argc, argv);
	std::cerr << "kwage_tools make_bloom|build_db|merge_db ..." << std::endl;
	return EXIT_FAILURE;
}
