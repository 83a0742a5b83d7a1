/* Test-infrastructure stub of the slice of the NCBI NGS C++ API that the reference's
 * make_bloom.cpp touches (make_bloom.cpp:170-300).  Reads come from a plain text file
 * "$KWAGE_READS_DIR/<accession>.reads" (one read per line, default directory "."), the
 * collection reports zero alignments so the reference takes its getReadRange(1,n,all) branch
 * (make_bloom.cpp:260-300), and every read has exactly one fragment. */
#ifndef KWAGE_ORACLE_STUB_NGS_HPP
#define KWAGE_ORACLE_STUB_NGS_HPP
#include <cassert>   // the real NGS headers pull in assert(), which make_bloom.cpp:98 relies on
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

namespace ngs {

class StringRef {
public:
	StringRef() : ptr(nullptr), len(0) {}
	StringRef(const char* p, size_t n) : ptr(p), len(n) {}
	const char* data() const { return ptr; }
	size_t size() const { return len; }
	std::string toString() const { return std::string(ptr, len); }
private:
	const char* ptr;
	size_t len;
};

struct Alignment { enum AlignmentCategory { primaryAlignment = 1, secondaryAlignment = 2, all = 3 }; };
struct Read { enum ReadCategory { fullyAligned = 1, partiallyAligned = 2, aligned = 3, unaligned = 4, all = 7 }; };

typedef std::shared_ptr< std::vector<std::string> > ReadStore;

class AlignmentIterator {
public:
	bool nextAlignment() { return false; }
	StringRef getAlignedFragmentBases() { return StringRef(); }
};

class ReadIterator {
public:
	ReadIterator() : cursor(0), started(false), fragment_done(true) {}
	explicit ReadIterator(const ReadStore& s) : store(s), cursor(0), started(false), fragment_done(true) {}
	bool nextRead()
	{
		if(!store) return false;
		if(started) ++cursor;
		started = true;
		fragment_done = false;
		return cursor < store->size();
	}
	bool nextFragment()
	{
		if(fragment_done) return false;
		fragment_done = true;
		return true;
	}
	StringRef getFragmentBases() { return StringRef((*store)[cursor].data(), (*store)[cursor].size()); }
private:
	ReadStore store;
	size_t cursor;
	bool started, fragment_done;
};

class ReadCollection {
public:
	explicit ReadCollection(const ReadStore& s) : store(s) {}
	uint64_t getAlignmentCount(Alignment::AlignmentCategory) { return 0; }
	AlignmentIterator getAlignments(Alignment::AlignmentCategory) { return AlignmentIterator(); }
	uint64_t getReadCount(Read::ReadCategory c) { return (c == Read::unaligned) ? 0 : store->size(); }
	ReadIterator getReads(Read::ReadCategory) { return ReadIterator(store); }
	ReadIterator getReadRange(uint64_t, uint64_t, Read::ReadCategory) { return ReadIterator(store); }
private:
	ReadStore store;
};

} // namespace ngs

namespace ncbi {
struct NGS {
	static ngs::ReadCollection openReadCollection(const std::string& accession)
	{
		const char* dir = std::getenv("KWAGE_READS_DIR");
		const std::string path = std::string(dir ? dir : ".") + "/" + accession + ".reads";
		std::ifstream fin(path.c_str());
		if(!fin) throw "NGS stub: unable to open reads file";
		ngs::ReadStore store(new std::vector<std::string>());
		std::string line;
		while(std::getline(fin, line)) store->push_back(line); // empty lines are zero-length reads
		return ngs::ReadCollection(store);
	}
};
} // namespace ncbi
#endif
