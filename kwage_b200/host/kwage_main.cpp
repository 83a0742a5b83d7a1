// kwage -- search KWAGE bit-sliced databases with DNA queries.  Same command line as the reference's
// kwage (options.cpp:39-192: -o, --o.csv, --o.json, -t, -d, -i, positional sequences) plus
// --device <n>.  Output assembly follows kwage.cpp:189-319 (results sorted by num_kmers_found,
// command-line sequences first, then file records by id).
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <sstream>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <zlib.h>

#include "kwage_host.h"

using namespace kwage;

namespace {

bool ends_with(const std::string& s, const std::string& suffix)
{
	return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// -d accepts files or directories searched recursively for *.db (options.cpp:30-33, file_util.h)
void collect_db_files(const std::string& path, std::deque<std::string>& out)
{
	struct stat st;
	if (::stat(path.c_str(), &st) != 0) throw __FILE__ ":main: Unable to stat database path";
	if (S_ISDIR(st.st_mode)) {
		DIR* d = opendir(path.c_str());
		if (!d) return;
		std::vector<std::string> names;
		while (struct dirent* e = readdir(d)) {
			const std::string n = e->d_name;
			if (n != "." && n != "..") names.push_back(n);
		}
		closedir(d);
		std::sort(names.begin(), names.end());
		for (size_t i = 0; i < names.size(); ++i) {
			const std::string child = path + "/" + names[i];
			struct stat cs;
			if (::stat(child.c_str(), &cs) != 0) continue;
			if (S_ISDIR(cs.st_mode)) collect_db_files(child, out);
			else if (ends_with(child, ".db")) out.push_back(child);
		}
	} else if (ends_with(path, ".db")) {
		out.push_back(path);
	}
}

// FASTA / FASTQ records (gz aware), bases upper-cased, white space dropped, records without
// bases skipped (parse_sequence.cpp:72-262)
bool read_queries(const std::string& path, std::vector<std::string>& names, std::vector<std::string>& seqs)
{
	gzFile gz = gzopen(path.c_str(), "rb");
	if (!gz) return false;
	gzbuffer(gz, 1 << 20);
	std::string line, name, seq;
	char buf[1 << 16];
	bool fastq = false, first = true, have = false;
	int fq_line = 0;
	for (;;) {
		line.clear();
		bool any = false;
		while (gzgets(gz, buf, sizeof(buf))) {
			any = true;
			size_t n = std::strlen(buf);
			const bool eol = n && buf[n - 1] == '\n';
			while (n && (buf[n - 1] == '\n' || buf[n - 1] == '\r')) --n;
			line.append(buf, n);
			if (eol) break;
		}
		if (!any) break;
		if (first && !line.empty()) { fastq = line[0] == '@'; first = false; }
		if (fastq) {
			if (fq_line == 0) { name = line; while (!name.empty() && (name[0] == '@' || std::isspace((unsigned char)name[0]))) name.erase(0, 1); }
			else if (fq_line == 1) {
				seq.clear();
				for (size_t i = 0; i < line.size(); ++i) if (!std::isspace((unsigned char)line[i])) seq.push_back((char)std::toupper((unsigned char)line[i]));
				if (!seq.empty()) { names.push_back(name); seqs.push_back(seq); }
			}
			fq_line = (fq_line + 1) & 3;
			continue;
		}
		if (line.find('>') != std::string::npos) {
			if (have && !seq.empty()) { names.push_back(name); seqs.push_back(seq); seq.clear(); }
			name = line;
			while (!name.empty() && (name[0] == '>' || std::isspace((unsigned char)name[0]))) name.erase(0, 1);
			have = true;
		} else {
			for (size_t i = 0; i < line.size(); ++i) if (!std::isspace((unsigned char)line[i])) seq.push_back((char)std::toupper((unsigned char)line[i]));
		}
	}
	if (!fastq && !seq.empty()) { names.push_back(name); seqs.push_back(seq); }
	gzclose(gz);
	return true;
}

int usage()
{
	std::cerr << "Usage for kwage (B200):\n"
		"\t[-o <output file>] (default is stdout)\n"
		"\t[--o.csv (output CSV) | --o.json (output JSON)]\n"
		"\t[-t <search threshold>] (default is 1)\n"
		"\t-d <database search path> (can be repeated)\n"
		"\t[-i <input sequence file>] (can be repeated)\n"
		"\t[--device <CUDA device>] (default is 0; can be repeated or a comma separated list: the database files are\n"
		"\t\tshared out over the devices, one host thread per device, like the reference's OpenMP loop over files)\n"
		"\t[--max-slab-gib <GiB>] (database files with equal Bloom parameters share one slab in HBM up to this size;\n"
		"\t\t0 = one file at a time; default is 48)\n"
		"\t[<DNA sequence>] (can be repeated)\n";
	return EXIT_FAILURE;
}

} // namespace

int main(int argc, char* argv[])
{
	try {
		SearchOptions opt;
		std::string output_file;
		std::deque<std::string> db_paths, query_files, query_seq;
		double max_slab_gib = 48.0;
		std::vector<int> devices;
		for (int i = 1; i < argc; ++i) {
			const std::string a = argv[i];
			if (a == "-o" && i + 1 < argc) output_file = argv[++i];
			else if (a == "--o.csv") opt.output_format = SearchOptions::OUTPUT_CSV;
			else if (a == "--o.json") opt.output_format = SearchOptions::OUTPUT_JSON;
			else if (a == "-t" && i + 1 < argc) opt.threshold = (float)std::atof(argv[++i]);
			else if (a == "-d" && i + 1 < argc) db_paths.push_back(argv[++i]);
			else if (a == "-i" && i + 1 < argc) query_files.push_back(argv[++i]);
			else if (a == "--device" && i + 1 < argc) {
				std::stringstream list(argv[++i]);
				std::string item;
				while (std::getline(list, item, ',')) if (!item.empty()) devices.push_back(std::atoi(item.c_str()));
			}
			else if (a == "--max-slab-gib" && i + 1 < argc) max_slab_gib = std::atof(argv[++i]);
			else if (a == "-h" || a == "-?") return usage();
			else if (!a.empty() && a[0] == '-') { std::cerr << '"' << a << "\" is not a valid option!" << std::endl; return usage(); }
			else query_seq.push_back(a);
		}
		if (db_paths.empty()) { std::cerr << "Please provide at least one database file to search (-d)" << std::endl; return usage(); }
		if (query_files.empty() && query_seq.empty()) { std::cerr << "Please provide at least one query sequence or file" << std::endl; return usage(); }
		if (!(opt.threshold > 0.0f && opt.threshold <= 1.0f)) { std::cerr << "Please provide: 0.0 < search threshold <= 1.0" << std::endl; return usage(); }

		std::deque<std::string> subject_files;
		for (size_t i = 0; i < db_paths.size(); ++i) collect_db_files(db_paths[i], subject_files);
		if (subject_files.empty()) { std::cerr << "Did not find any database files to search" << std::endl; return EXIT_FAILURE; }

		std::vector<std::string> file_names, file_seqs;
		for (size_t i = 0; i < query_files.size(); ++i)
			if (!read_queries(query_files[i], file_names, file_seqs)) throw __FILE__ ":main: Unable to open query file";
		std::vector<std::string> cmd_seqs(query_seq.begin(), query_seq.end());
		std::vector<size_t> cmd_ids(cmd_seqs.size()), file_ids(file_seqs.size());
		for (size_t i = 0; i < cmd_ids.size(); ++i) cmd_ids[i] = i;
		for (size_t i = 0; i < file_ids.size(); ++i) file_ids[i] = i;

		std::unordered_map<size_t, std::deque<MatchResult> > cmd_results, file_results;
		// consecutive files with equal Bloom parameters share one column slab in HBM (one pass of wide rows instead of
		// one pass of 256-byte rows per file); the match set is the same either way
		const uint64_t max_slab_bytes = (uint64_t)(max_slab_gib * 1073741824.0);
		std::vector<std::vector<std::string> > groups;
		for (size_t f = 0; f < subject_files.size();) {
			std::vector<std::string> group(1, subject_files[f]);
			uint64_t bytes = SubjectDatabase::slab_bytes(subject_files[f]);
			size_t g = f + 1;
			while (g < subject_files.size() && SubjectDatabase::compatible(subject_files[f], subject_files[g]) &&
			       bytes + SubjectDatabase::slab_bytes(subject_files[g]) <= max_slab_bytes) {
				bytes += SubjectDatabase::slab_bytes(subject_files[g]);
				group.push_back(subject_files[g++]);
			}
			groups.push_back(group);
			f = g;
		}
		if (devices.empty()) devices.push_back(opt.device);
		// Several devices: rounds of one slab per device.  Every device searches all the queries against its slab and the
		// per-slab hit lists meet on the first device through ONE NCCL exchange inside the library (kwg_search_gather) --
		// what the reference's critical section does with its thread-local maps (kwage.cpp:154-177).  Slabs left over
		// (fewer than devices) and single-device runs take the plain path below.
		size_t first_plain_group = 0;
		bool distinct = true;
		for (size_t a = 0; a < devices.size(); ++a)
			for (size_t b = a + 1; b < devices.size(); ++b) distinct = distinct && devices[a] != devices[b];
		if (devices.size() > 1 && distinct && groups.size() >= devices.size()) {
			const size_t n = devices.size();
			std::vector<kwg_comm_t*> comms(n, (kwg_comm_t*)NULL);
			// NCCL prints its version banner (NCCL_DEBUG=VERSION) on stdout while the communicators are created, and
			// stdout carries the results: it points at stderr for the duration of that call
			std::cout.flush();
			fflush(stdout);
			const int saved_stdout = dup(1);
			if (saved_stdout >= 0) dup2(2, 1);
			const int comm_rc = kwg_comm_create_all(comms.data(), (int)n, devices.data());
			fflush(stdout);
			if (saved_stdout >= 0) { dup2(saved_stdout, 1); close(saved_stdout); }
			if (comm_rc == KWG_OK) {
				std::vector<std::string> all_seqs(cmd_seqs);
				all_seqs.insert(all_seqs.end(), file_seqs.begin(), file_seqs.end());
				const size_t full_rounds = groups.size() / n;
				std::string round_error;
				for (size_t r = 0; r < full_rounds && round_error.empty(); ++r) {
					std::vector<SubjectDatabase*> subj(n, (SubjectDatabase*)NULL);
					std::vector<std::string> err(n);
					std::vector<std::thread> th;
					for (size_t w = 0; w < n; ++w)
						th.push_back(std::thread([&, w]() {
							try { subj[w] = new SubjectDatabase(groups[r * n + w], devices[w]); }
							catch (const char* e) { err[w] = e; } catch (...) { err[w] = "unhandled error"; }
						}));
					for (size_t w = 0; w < n; ++w) th[w].join();
					th.clear();
					bool ok = true;
					for (size_t w = 0; w < n; ++w) ok = ok && err[w].empty();
					std::vector<uint32_t> filter0(n + 1, 0);
					for (size_t w = 0; w < n && ok; ++w) filter0[w + 1] = filter0[w] + subj[w]->header().num_filter;
					std::vector<kwg_hit_t> hits;
					std::vector<uint32_t> nk;
					if (ok) {
						for (size_t w = 0; w < n; ++w)
							th.push_back(std::thread([&, w]() {
								try {
									std::vector<kwg_hit_t> h;
									std::vector<uint32_t> k;
									subj[w]->search_gather(comms[w], 0, filter0[w], all_seqs, opt.threshold, h, k);
									if (w == 0) { hits.swap(h); nk.swap(k); }
								}
								catch (const char* e) { err[w] = e; } catch (...) { err[w] = "unhandled error"; }
							}));
						for (size_t w = 0; w < n; ++w) th[w].join();
						for (size_t w = 0; w < n; ++w) ok = ok && err[w].empty();
					}
					if (ok) {
						for (size_t i = 0; i < hits.size(); ++i) {
							size_t w = 0;
							while (w + 1 < n && hits[i].filter >= filter0[w + 1]) ++w;
							const FilterInfo info = subj[w]->filter_info(hits[i].filter - filter0[w]);
							const size_t q = hits[i].query;
							if (q < cmd_seqs.size()) cmd_results[q].push_back(MatchResult(hits[i].num_match, nk[q], info));
							else file_results[q - cmd_seqs.size()].push_back(MatchResult(hits[i].num_match, nk[q], info));
						}
					}
					for (size_t w = 0; w < n; ++w) { if (!err[w].empty() && round_error.empty()) round_error = err[w]; delete subj[w]; }
				}
				for (size_t w = 0; w < n; ++w) kwg_comm_destroy(comms[w]);
				if (!round_error.empty()) { std::cerr << "Caught the error " << round_error << std::endl; return EXIT_FAILURE; }
				first_plain_group = full_rounds * n;
			}
			else std::cerr << "kwage: NCCL is not available (" << kwg_last_error() << "); merging the hit lists on the host" << std::endl;
		}
		// one host thread per device takes slabs off a shared counter (the reference: one OpenMP thread per file,
		// kwage.cpp:76-87); every thread collects its matches privately and they are merged afterwards
		const size_t n_plain = groups.size() - first_plain_group;
		const size_t n_workers = std::min(devices.size(), n_plain);
		std::vector<std::unordered_map<size_t, std::deque<MatchResult> > > w_cmd(n_workers), w_file(n_workers);
		std::vector<std::string> w_error(n_workers);
		std::atomic<size_t> next(first_plain_group);
		std::vector<std::thread> workers;
		for (size_t w = 0; w < n_workers; ++w) {
			workers.push_back(std::thread([&, w]() {
				try {
					SearchOptions o = opt;
					o.device = devices[w];
					for (size_t gi = next.fetch_add(1); gi < groups.size(); gi = next.fetch_add(1)) {
						SubjectDatabase subject(groups[gi], o.device);
						subject.search(w_cmd[w], cmd_seqs, cmd_ids, o);
						subject.search(w_file[w], file_seqs, file_ids, o);
					}
				}
				catch (const char* error) { w_error[w] = error; }
				catch (const std::exception& error) { w_error[w] = error.what(); }
				catch (...) { w_error[w] = "unhandled error"; }
			}));
		}
		for (size_t w = 0; w < workers.size(); ++w) workers[w].join();
		for (size_t w = 0; w < n_workers; ++w) {
			if (!w_error[w].empty()) { std::cerr << "Caught the error " << w_error[w] << std::endl; return EXIT_FAILURE; }
			for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = w_cmd[w].begin(); i != w_cmd[w].end(); ++i)
				cmd_results[i->first].insert(cmd_results[i->first].end(), i->second.begin(), i->second.end());
			for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = w_file[w].begin(); i != w_file[w].end(); ++i)
				file_results[i->first].insert(file_results[i->first].end(), i->second.begin(), i->second.end());
		}
		for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = cmd_results.begin(); i != cmd_results.end(); ++i)
			std::sort(i->second.begin(), i->second.end());
		for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = file_results.begin(); i != file_results.end(); ++i)
			std::sort(i->second.begin(), i->second.end());

		std::ofstream fout;
		if (!output_file.empty()) {
			fout.open(output_file.c_str());
			if (!fout) throw __FILE__ ":main: Unable to open output file";
		}
		std::ostream& out = output_file.empty() ? std::cout : fout;
		const bool multiple = (cmd_results.size() + file_results.size()) > 1;
		if (opt.output_format == SearchOptions::OUTPUT_CSV) write_csv_header(out); else write_json_header(out, multiple);
		bool first = true;
		std::vector<size_t> ids;
		for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = cmd_results.begin(); i != cmd_results.end(); ++i) ids.push_back(i->first);
		std::sort(ids.begin(), ids.end());
		for (size_t i = 0; i < ids.size(); ++i) {
			std::stringstream name;
			name << "command line seq " << ids[i];
			if (opt.output_format == SearchOptions::OUTPUT_CSV) write_csv(out, name.str(), cmd_results[ids[i]]);
			else write_json(out, name.str(), multiple, first, opt.threshold, cmd_results[ids[i]]);
			first = false;
		}
		ids.clear();
		for (std::unordered_map<size_t, std::deque<MatchResult> >::iterator i = file_results.begin(); i != file_results.end(); ++i) ids.push_back(i->first);
		std::sort(ids.begin(), ids.end());
		for (size_t i = 0; i < ids.size(); ++i) {
			if (opt.output_format == SearchOptions::OUTPUT_CSV) write_csv(out, file_names[ids[i]], file_results[ids[i]]);
			else write_json(out, file_names[ids[i]], multiple, first, opt.threshold, file_results[ids[i]]);
			first = false;
		}
		if (opt.output_format == SearchOptions::OUTPUT_JSON) write_json_footer(out, multiple);
	}
	catch (const char* error) {
		std::cerr << "Caught the error " << error << std::endl;
		return EXIT_FAILURE;
	}
	catch (const std::exception& error) {
		std::cerr << "Caught the error " << error.what() << std::endl;
		return EXIT_FAILURE;
	}
	catch (...) {
		std::cerr << "Caught an unhandled error" << std::endl;
		return EXIT_FAILURE;
	}
	return EXIT_SUCCESS;
}
