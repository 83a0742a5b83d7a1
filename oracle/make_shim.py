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


def main():
    ref, out = sys.argv[1], sys.argv[2]
    text = open(ref + "/build_db.cpp").read()
    a, b = text.index(BEGIN), text.index(END)
    assert a < b and text.count(BEGIN) == 1 and text.count(END) == 1, "the reference's chunk loop was not found where expected"
    patched = '#include <vector>\n#include "kwage_cuda.h"\n' + text[:a] + SHIM + text[b:]
    open(out, "w").write(patched)


if __name__ == "__main__":
    main()
