// kwage_tools -- make_bloom / build_db front ends of the host layer (the roles of the reference's
// bff.cpp test rig and of a maestro worker's two work items, without MPI):
//   kwage_tools make_bloom [options] <accession | reads file>...     -> <bloom dir>/<ACC>.bloom
//   kwage_tools build_db -o <out.db> <a.bloom> <b.bloom> ...          (parameters come from the first file)
// Options of make_bloom mirror maestro's (options.cpp:404-820): -k, --min-kmer-count, -p,
// --len.min, --len.max, --bloom <dir>; plus --device and --num-bases (metadata override).
#include <cstdlib>
#include <cstring>
#include <chrono>

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

int main(int argc, char** argv)
{
	if (argc >= 2 && std::strcmp(argv[1], "make_bloom") == 0) return cmd_make_bloom(argc, argv);
	if (argc >= 2 && std::strcmp(argv[1], "build_db") == 0) return cmd_build_db(argc, argv);
	std::cerr << "kwage_tools make_bloom|build_db ..." << std::endl;
	return EXIT_FAILURE;
}
