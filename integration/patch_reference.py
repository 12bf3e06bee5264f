#!/usr/bin/env python
"""Apply the libpbk seam to a COPY of the reference's counter.h.

    python integration/patch_reference.py <reference dir> <output dir> [--emit-diff counter_h.patch]

Copies the reference's top-level sources (*.cpp, *.h) into <output dir> and rewrites three member-function bodies of
Counter<KMER> in the copy so that they forward to pbk::Counter (platanus_b_b200/host/pbk_counter.hpp, the C ABI of
include/pbk.h underneath):

    makeKmerReadDistributionMT                         counter.h:276-383   first k: reads -> kmerFP, distributions
    makeKmerReadDistributionConsideringPreviousGraph   counter.h:663-750   later k: contig-seeded table + reads
    pickupReadMatchedEdgeKmer                          counter.h:870-910   which reads the next round keeps

plus one #include and one data member.  Nothing else of the reference changes: graph.h, assemble.cpp, kmer_divide.cpp ...
compile and link unchanged against the patched header (`make -C oracle ref_patched`).  The replacement bodies below are
ours; the reference text is only located (by its function signatures), never stored here -- which is why this is a
script and not a .patch file (a unified diff would carry the removed reference lines; --emit-diff writes one next to
the patched copy for whoever wants to read or `git apply` it)."""
import argparse
import difflib
import glob
import os
import shutil
import sys

INCLUDE = '#include "platanus_b_b200/host/pbk_reference_shim.hpp"   // B200 k-mer counter behind a C ABI (include/pbk.h)\n'
MEMBER = "    std::shared_ptr<pbk::Counter> gpu;   // the device-side counter of this object (created on first use)\n"

BODY_MT = r'''{
    // B200 path: the reads of the numThread SEQ temp files are counted by libpbk (hash table in HBM); what this function
    // leaves behind is unchanged -- kmerFP, occurrenceDistribution, lengthDistribution, maxOccurrence, the return value.
    this->kmerLength = kLength;
    if (kmerFP != NULL)
        fclose(kmerFP);
    kmerFP = platanus::makeTemporaryFile();
    try {
        if (!gpu) gpu.reset(new pbk::Counter(kLength));
        const unsigned long long doubleHashSize = gpu->makeKmerReadDistributionMT(kLength, readFP, memory, numThread);
        lengthDistribution.assign(platanus::ConstParam::MAX_READ_LEN + 1, 0);
        for (unsigned long long i = 0; i < platanus::ConstParam::MAX_READ_LEN + 1; ++i)
            lengthDistribution[i] = gpu->getLengthDistributionI(i);
        pbk::shim::publish(*gpu, kLength, kmerFP, occurrenceDistribution, this->maxOccurrence);
        return doubleHashSize;
    } catch (const pbk::ErrorBase &e) {
        pbk::shim::rethrow(e);
    }
    return 0;
}



'''

BODY_PREV = r'''{
    KMER kmer(k);
    std::cerr << "K = " << k << ", saving additional kmers(not found in contigs) from reads..." << std::endl;
    if (kmerFP != NULL)
        fclose(kmerFP);
    kmerFP = platanus::makeTemporaryFile();
    // B200 path: the k-mers already in occurrenceTable (seeded from the previous round's contigs) keep their values, every
    // other k-mer of the reads is counted (pbk_seed_entries + pbk_push_reads); the host table is emptied like the original's dump.
    try {
        std::vector<uint64_t> words;
        std::vector<uint16_t> values;
        pbk::shim::tableEntries(kmer, occurrenceTable, true, words, values);
        if (!gpu) gpu.reset(new pbk::Counter(k));
        gpu->makeKmerReadDistributionSeeded(k, readFP, memory, numThread, words.data(), values.data(), values.size());
        pbk::shim::publish(*gpu, k, kmerFP, occurrenceDistribution, this->maxOccurrence);
    } catch (const pbk::ErrorBase &e) {
        pbk::shim::rethrow(e);
    }
    return 0;
}

'''

BODY_PICKUP = r'''{
    FILE *newReadFP = platanus::makeTemporaryFile();
    // B200 path: the edge k-mers of occurrenceTable go to the device, then every read of this file is probed in one pass
    // (pbk_match_reads); kept records are copied verbatim.  Called from an OpenMP loop (assemble.cpp:431-433): one device
    // context per Counter, so the calls take turns.
    pbk::shim::SeqFile sf;
    sf.read(*readFP);
    const size_t nReads = sf.offsets.size() - 1;
    std::vector<uint8_t> matched(nReads, 0);
    #pragma omp critical (pbk_gpu_counter)
    {
        try {
            KMER kmer(kmerLength);
            std::vector<uint64_t> words;
            std::vector<uint16_t> values;
            pbk::shim::tableEntries(kmer, occurrenceTable, false, words, values);
            if (!gpu) gpu.reset(new pbk::Counter(kmerLength));
            gpu->beginCounting(kmerLength);
            gpu->loadEntries(words.data(), values.data(), values.size());
            if (nReads) gpu->matchReadsPlatanus(sf.bases.data(), sf.offsets.data(), nReads, sf.npos.data(), sf.nposOffsets.data(), matched.data());
        } catch (const pbk::ErrorBase &e) {
            e.showErrorMessage();
            exit(platanus::DOUBLEHASH);             // exceptions must not leave an OpenMP region
        }
    }
    std::vector<char> record;
    for (size_t r = 0; r < nReads; ++r) {
        if (!matched[r]) continue;
        const long n = sf.recordStart[r + 1] - sf.recordStart[r];
        record.resize(n);
        fseek(*readFP, sf.recordStart[r], SEEK_SET);
        if (fread(record.data(), 1, n, *readFP) != static_cast<size_t>(n)) throw platanus::ReadError();
        fwrite(record.data(), 1, n, newReadFP);
    }
    *readFP = newReadFP;
}


'''

# (signature that opens the function, text that starts whatever follows the function) -- anchors only
SEAMS = [
    ("unsigned long long Counter<KMER>::makeKmerReadDistributionMT(",
     "//////////////////////////////////////////////////////////////////////////////////////\n// count kmer multi thread for initial", BODY_MT),
    ("unsigned long long Counter<KMER>::makeKmerReadDistributionConsideringPreviousGraph(",
     "//////////////////////////////////////////////////////////////////////////////////////\n// count kmer multi thread\n", BODY_PREV),
    ("void Counter<KMER>::pickupReadMatchedEdgeKmer(FILE **readFP)",
     "//////////////////////////////////////////////////////////////////////////////////////\n// sort kmer", BODY_PICKUP),
]


def patch_counter_h(text: str) -> str:
    assert "pbk_reference_shim.hpp" not in text, "already patched"
    anchor = "#include <memory>\n"
    assert text.count(anchor) == 1
    text = text.replace(anchor, anchor + INCLUDE)
    anchor = "    u64_t maxOccurrence;\n"
    assert text.count(anchor) == 1
    text = text.replace(anchor, anchor + MEMBER)
    for sig, nxt, body in SEAMS:
        assert text.count(sig) == 1, sig
        a = text.index(sig)
        b = text.index("{", a)
        e = text.index(nxt, b)
        text = text[:b] + body + text[e:]
    return text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("out")
    ap.add_argument("--emit-diff", default="")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for f in glob.glob(os.path.join(a.reference, "*.cpp")) + glob.glob(os.path.join(a.reference, "*.h")):
        shutil.copy(f, a.out)
    src = os.path.join(a.out, "counter.h")
    orig = open(src).read()
    new = patch_counter_h(orig)
    open(src, "w").write(new)
    if a.emit_diff:
        open(a.emit_diff, "w").writelines(difflib.unified_diff(orig.splitlines(True), new.splitlines(True), "a/counter.h", "b/counter.h"))
    print(f"patched {src}: {len(SEAMS)} member functions forward to pbk::Counter", file=sys.stderr)


if __name__ == "__main__":
    main()
