#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Proves the drop-in claim of INTEGRATION.md section 2 against the reference's own source: reads
build_db.cpp where it lies under /root/reference, replaces the body of the chunk loop (the per-filter read + bit-by-bit
scatter, build_db.cpp:269-304) by the kwg_transpose shim, and writes the patched translation unit to oracle/_build/
(git-ignored -- no reference source enters the repository).  oracle/Makefile then compiles it with the reference's other
sources and links the result against libkwage_cuda.so: oracle/_ref/ref_driver_gpu is the UNMODIFIED reference host code
calling the CUDA library for the transposition.

  python oracle/make_shim.py /root/reference oracle/_build/build_db_gpu.cpp
"""
import sys

BEGIN = "BitVector src(num_buffer_slice);"
END = "// Update the crc32 value with the contents of the current destination buffer"

SHIM = r'''// ---- kwage_b200 shim (INTEGRATION.md section 2): the chunk of every filter is read into one staging buffer, the
			// bitwise transposition runs on the GPU; crc bookkeeping, file writing and everything else is the reference's
			std::vector<unsigned char> kwg_src_all(num_filter * (num_buffer_slice / 8));
			std::vector<const uint8_t*> kwg_chunk(num_filter);

			for(size_t j = 0;j < num_filter;++j){

				unsigned char* p = &kwg_src_all[j * (num_buffer_slice / 8)];

				fin_ptr[j]->read( (char*)p, num_buffer_slice / 8 );

				if( !(*fin_ptr[j]) ){
					throw __FILE__ ":build_db: Error reading filter bytes";
				}

				filter_crc32[j].second = ::crc32_z(filter_crc32[j].second, p, num_buffer_slice / 8);
				kwg_chunk[j] = p;
			}

			if( kwg_transpose(0, kwg_chunk.data(), (uint32_t)num_filter, num_buffer_slice, dest) != KWG_OK ){
				cerr << "kwg_transpose: " << kwg_last_error() << endl;
				throw __FILE__ ":build_db: kwg_transpose failed";
			}

			'''


# ---- make_bloom.cpp: count_words() buffers the fragments and hands whole batches to
# kwg_bloom_add_reads (stream order kept), the number of valid k-mers comes back from the device, the fold of valid_bits
# is replaced by kwg_bloom_finalize.  Everything else -- NGS iteration, the max_num_kmer checks, optimal_bloom_param,
# update_crc32, set_info, binary_write, the status codes -- is the reference's.
MB_COUNT_WORDS = "void count_words(CountingBloom *m_count_ptr, vector<BitVector> &m_valid_bits, \n\tsize_t &m_num_valid_kmer, const ngs::StringRef &m_seq, \n\tconst size_t &m_hash_seq_mask, const size_t &m_hash_count_mask, \n\tconst MaestroOptions &m_opt)\n{"
MB_AFTER_ALLOC = "memset(bcount, 0, num_count_bloom);"
MB_BEFORE_PARAM = "m_param = optimal_bloom_param(m_opt.kmer_len,"
MB_FOLD_BEGIN = "BitVector::BLOCK *dst_ptr = filter.ptr();"
MB_FOLD_END = "#ifdef DEBUG_BLOOM\n\t\tcerr << \"[\" << mpi_rank << \"] Set Bloom filter bits\" << endl;"

MB_PRELUDE = r'''#include <string>
#include <vector>
#include "kwage_cuda.h"
// ---- kwage_b200 shim state: one construction handle per make_bloom_filter() call (one per MPI worker rank)
static kwg_bloom_t* kwg_handle = NULL;
static std::string kwg_flat;
static std::vector<uint64_t> kwg_offsets(1, 0);
static void kwg_flush(size_t &m_num_valid_kmer)
{
	if(kwg_offsets.size() > 1){
		if(kwg_bloom_add_reads(kwg_handle, kwg_flat.data(), kwg_offsets.data(), kwg_offsets.size() - 1) != KWG_OK){
			throw __FILE__ ":make_bloom_filter: kwg_bloom_add_reads failed";
		}
		kwg_flat.clear();
		kwg_offsets.assign(1, 0);
	}
	uint64_t n = 0;
	if(kwg_bloom_num_valid(kwg_handle, &n) != KWG_OK){
		throw __FILE__ ":make_bloom_filter: kwg_bloom_num_valid failed";
	}
	m_num_valid_kmer = n;
}
'''

MB_COUNT_BODY = r'''
	// kwage_b200 shim: the fragment joins the current batch; batches of 64 MiB go to the device in stream order
	kwg_flat.append( m_seq.data(), m_seq.size() );
	kwg_offsets.push_back( kwg_flat.size() );
	if( kwg_flat.size() >= (64u << 20) ){
		kwg_flush(m_num_valid_kmer);
	}
}
#if 0
{'''


def patch_make_bloom(ref, out):
    text = open(ref + "/make_bloom.cpp").read()
    for anchor in (MB_COUNT_WORDS, MB_AFTER_ALLOC, MB_BEFORE_PARAM, MB_FOLD_BEGIN, MB_FOLD_END):
        assert text.count(anchor) == 1, "anchor not found exactly once in the reference's make_bloom.cpp: " + anchor[:50]
    # the counting filters live on the device: a handle right after the reference allocated its own tables
    text = text.replace(MB_AFTER_ALLOC, MB_AFTER_ALLOC + """
		if(kwg_handle != NULL){ kwg_bloom_destroy(kwg_handle); kwg_handle = NULL; }
		kwg_flat.clear(); kwg_offsets.assign(1, 0);
		if(kwg_bloom_create(&kwg_handle, 0, m_opt.kmer_len, m_opt.min_kmer_count,
			m_progress.log_2_counting_filter_len, m_opt.max_log_2_filter_len) != KWG_OK){
			cerr << "kwg_bloom_create: " << kwg_last_error() << endl;
			throw __FILE__ ":make_bloom_filter: kwg_bloom_create failed";
		}
""")
    # the last batch and the final count, before the parameters are chosen
    # (placed in front of the try block around optimal_bloom_param: its catch(...) means "no valid parameters", not "device error")
    at = text.rindex("try{", 0, text.index(MB_BEFORE_PARAM))
    text = text[:at] + "kwg_flush(m_progress.num_kmer);\n\n\t\t" + text[at:]
    # the fold of valid_bits[h] into the filter
    a, b = text.index(MB_FOLD_BEGIN), text.index(MB_FOLD_END)
    text = text[:a] + """if(kwg_bloom_finalize(kwg_handle, m_param.log_2_filter_len, m_param.num_hash, filter.ptr()) != KWG_OK){
			throw __FILE__ ":make_bloom_filter: kwg_bloom_finalize failed";
		}
		kwg_bloom_destroy(kwg_handle);
		kwg_handle = NULL;

		""" + text[b:]
    # count_words: buffer instead of count (the reference's body is compiled out)
    text = text.replace(MB_COUNT_WORDS, MB_COUNT_WORDS + MB_COUNT_BODY) + "\n#endif\n"
    open(out, "w").write(MB_PRELUDE + text)


# ---- kwage.cpp: the body of search() -- k-mer extraction, the per-k-mer slice reads, AND /
# count loops and the match decision -- becomes one kwg_search call against the file's slices resident in HBM (loaded once
# per database file and thread); option parsing, FASTA iteration, the FilterInfo look-up, MatchResult, sorting and the
# CSV / JSON writers are the reference's.
KW_SIGNATURE = "const SearchOptions &m_opt)\n{"

KW_BODY = r'''
	// ---- kwage_b200 shim: this thread keeps the slices of the database file it is working on resident in HBM
	static thread_local kwg_db_t* kwg_db = NULL;
	static thread_local unsigned long int kwg_key[4] = {0, 0, 0, 0};
	const unsigned long int key[4] = {(unsigned long int)m_header.crc32, m_info_start, (unsigned long int)m_header.num_filter,
		(unsigned long int)m_header.log_2_filter_len};

	if( (kwg_db == NULL) || (memcmp(key, kwg_key, sizeof(key)) != 0) ){

		if(kwg_db != NULL){
			kwg_db_unload(kwg_db);
			kwg_db = NULL;
		}

		const size_t num_slice = size_t(1) << m_header.log_2_filter_len;
		std::vector<unsigned char> slices(num_slice*m_slice_size);

		m_fsubject.clear();
		m_fsubject.seekg(m_bloom_start);
		m_fsubject.read( (char*)slices.data(), slices.size() );

		if(!m_fsubject){
			throw __FILE__ ":search: Error reading slice from file (1)";
		}

		if(kwg_db_load(&kwg_db, 0, slices.data(), m_header.kmer_len, m_header.num_hash, m_header.log_2_filter_len,
			m_header.num_filter, 0, m_header.num_filter) != KWG_OK){

			cerr << "kwg_db_load: " << kwg_last_error() << endl;
			throw __FILE__ ":search: kwg_db_load failed";
		}

		memcpy(kwg_key, key, sizeof(key));
	}

	const char* query_ptr = m_query.c_str();
	const uint64_t query_len = m_query.size();
	uint32_t num_query_kmer = 0;
	kwg_hit_t* hits = NULL;
	uint64_t num_hit = 0;

	if(kwg_search_ptrs(kwg_db, &query_ptr, &query_len, 1, m_opt.threshold, &num_query_kmer, &hits, &num_hit) != KWG_OK){
		throw __FILE__ ":search: kwg_search failed";
	}

	unordered_map< size_t, deque<MatchResult> >::iterator result_iter = m_search_results.end();

	for(uint64_t h = 0;h < num_hit;++h){

		// Read the Bloom filter info exactly like the reference does (kwage.cpp:505-515)
		m_fsubject.clear();
		m_fsubject.seekg( m_info_start + hits[h].filter*sizeof(unsigned long int) );

		unsigned long int info_loc;

		m_fsubject.read( (char*)&info_loc, sizeof(unsigned long int) );
		m_fsubject.seekg(info_loc);

		FilterInfo info;

		binary_read(m_fsubject, info);

		if( result_iter == m_search_results.end() ){

			result_iter = m_search_results.find(m_query_id);

			if( result_iter == m_search_results.end() ){
				result_iter = m_search_results.insert( make_pair(m_query_id, deque<MatchResult>() ) ).first;
			}
		}

		result_iter->second.push_back( MatchResult(hits[h].num_match, num_query_kmer, info) );
	}

	kwg_free_hits(hits);

	return (num_hit > 0);
}
'''


def patch_kwage(ref, out):
    text = open(ref + "/kwage.cpp").read()
    assert text.count(KW_SIGNATURE) == 1, "the definition of search() was not found where expected"
    at = text.index(KW_SIGNATURE) + len(KW_SIGNATURE)
    # search() is the last function of the file: everything after its opening brace is its body
    assert "\nbool " not in text[at:] and "\nint " not in text[at:] and "\nvoid " not in text[at:]
    open(out, "w").write('#include <string.h>\n#include <vector>\n#include "kwage_cuda.h"\n' + text[:at] + KW_BODY)


def main():
    ref, out = sys.argv[1], sys.argv[2]
    if out.endswith("make_bloom_gpu.cpp"):
        return patch_make_bloom(ref, out)
    if out.endswith("kwage_gpu.cpp"):
        return patch_kwage(ref, out)
    text = open(ref + "/build_db.cpp").read()
    a, b = text.index(BEGIN), text.index(END)
    assert a < b and text.count(BEGIN) == 1 and text.count(END) == 1, "the reference's chunk loop was not found where expected"
    patched = '#include <vector>\n#include "kwage_cuda.h"\n' + text[:a] + SHIM + text[b:]
    open(out, "w").write(patched)


if __name__ == "__main__":
    main()
