"""GPU: the whole path -- reads -> .bloom -> .db -> search output -- side by side with the UNMODIFIED reference compiled
into oracle/_ref (it travels to the GPU box), at the reference's default threshold (min_kmer_count 5) on ragged reads with
non-ACGT bytes and lower case.  Every file must be byte-identical and the search output must list the same matches."""
import os
import subprocess

import numpy as np
import pytest

from kwage_b200 import hostapi as H
from kwage_b200.host import build as hbuild
from oracle import oracle_py as O
import synth_cases as S
import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not O.have_ref(), reason="needs the compiled reference (oracle/_ref)")]

N_ACC, N_READS, GENOME = 6, 30000, 150000


def accession_reads(j):
    """~19x coverage of a private genome, every 7th read comes from a genome shared by all accessions (so queries hit
    several filters), ragged lengths 40..150, some N / lower case."""
    rng = np.random.default_rng(5000 + j)
    own = O.gen_reads(7000 + j, 0, 1, GENOME)
    shared = O.gen_reads(6999, 0, 1, 20000)
    reads = []
    for r in range(N_READS):
        g = shared if r % 7 == 0 else own
        n = int(rng.integers(40, 151))
        a = int(rng.integers(0, len(g) - n))
        reads.append(g[a: a + n])
    bases = np.concatenate(reads)
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in reads])]).astype(np.uint64)
    return S.mutate(bases, 5000 + j, n_rate=701, lower_rate=11), offsets


def sha_file(path):
    return util.sha256(np.fromfile(path, dtype=np.uint8))


def test_reads_to_search_output_equals_the_reference(tmp_path):
    hbuild.build()
    ours, ref, reads_dir = tmp_path / "ours", tmp_path / "ref", tmp_path / "reads"
    for d in (ours, ref, reads_dir):
        d.mkdir()
    k, c, p, lmin, lmax = 31, 5, 0.25, 18, 24
    params = set()
    for j in range(N_ACC):
        acc = util.fixture_accession(j)
        bases, offsets = accession_reads(j)
        S.write_reads_file(str(reads_dir / (acc + ".reads")), bases, offsets)
        num_bp = int(offsets[-1])
        r = H.make_bloom_file(acc, str(reads_dir / (acc + ".reads")), num_bp, str(ours), k=k, min_kmer_count=c, p=p, min_log2=lmin, max_log2=lmax)
        g = O.ref_make_bloom(acc, str(reads_dir), str(ref), k, c, p, lmin, lmax, num_bp)
        assert r["status"] == H.STATUS_BLOOM_SUCCESS == g["status"], (r, g)
        assert (r["num_kmer"], r["log2_len"], r["num_hash"]) == (g["num_kmer"], g["log_2_filter_len"], g["num_hash"])
        assert sha_file(str(ours / (acc + ".bloom"))) == sha_file(str(ref / (acc + ".bloom")))
        params.add((r["log2_len"], r["num_hash"]))
    assert len(params) == 1, params                     # one database needs one parameter set
    L, h = params.pop()
    # database
    files = [str(ours / (util.fixture_accession(j) + ".bloom")) for j in range(N_ACC)]
    assert H.build_db(str(ours / "all.db"), k, L, h, files)
    listing = tmp_path / "blooms.txt"
    listing.write_text("\n".join(str(ref / (util.fixture_accession(j) + ".bloom")) for j in range(N_ACC)) + "\n")
    O.ref_driver("build_db", str(ref / "all.db"), k, L, h, str(listing))
    assert sha_file(str(ours / "all.db")) == sha_file(str(ref / "all.db"))
    # search: pieces of three private genomes, of the shared genome, and noise
    fa = tmp_path / "q.fa"
    with open(fa, "w") as f:
        for j in (0, 3, 5):
            f.write(">own_%d\n%s\n" % (j, bytes(O.gen_reads(7000 + j, 0, 1, GENOME)[1000: 2200]).decode()))
        f.write(">shared\n%s\n" % bytes(O.gen_reads(6999, 0, 1, 20000)[500: 1500]).decode())
        f.write(">noise\n%s\n" % bytes(O.gen_reads(1, 0, 1, 900)).decode())
    for t in ("1", "0.5", "0.05"):
        a = subprocess.run([H.KWAGE_BIN, "-d", str(ours / "all.db"), "-i", str(fa), "-t", t, "--o.csv"], capture_output=True, text=True)
        b = O.ref_kwage(["-d", str(ref / "all.db"), "-i", str(fa), "-t", t, "--o.csv"], omp_threads=1)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert sorted(a.stdout.splitlines()) == sorted(b.stdout.splitlines()), t
    rows = [x for x in a.stdout.splitlines() if x.startswith('"shared"')]
    assert len(rows) == N_ACC                            # the shared genome is in every accession
