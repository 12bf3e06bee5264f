// pbk_reference_shim.hpp -- glue between the reference's own types and pbk::Counter, used ONLY by the patched
// reference (integration/counter_h.patch: the bodies of Counter<KMER>::makeKmerReadDistributionMT,
// makeKmerReadDistributionConsideringPreviousGraph and pickupReadMatchedEdgeKmer, counter.h:276-383, 663-750, 870-910).
// Included from counter.h after common.h / kmer.h / doubleHash.h, so platanus::, KMER and DoubleHash are visible here.
// Keys cross to the C ABI as the u64 words KMER::writeKey puts into kmerFP (kmer.h:119, 237) -- obtained with the
// reference's own serialiser through an in-memory FILE, so that no knowledge of Kmer31 / KmerN<BinstrNN> / binstr_t
// internals is needed on this side.
#ifndef PBK_REFERENCE_SHIM_HPP
#define PBK_REFERENCE_SHIM_HPP

#include "pbk_counter.hpp"

#include <memory>

namespace pbk {
namespace shim {

// pbk errors -> the reference's exception types, so that main.cpp:121-124 prints and exits as before.  A GPU failure has no
// counterpart in platanus::ERROR; it is reported as DOUBLEHASH (the table could not be built), exit code 10.
inline void rethrow(const pbk::ErrorBase &e)
{
    switch (e.getID()) {
    case pbk::E_READ: throw platanus::ReadError();
    case pbk::E_KMERDIST: throw platanus::KmerDistError();
    case pbk::E_TMP: throw platanus::TMPError();
    case pbk::E_FOPEN: throw platanus::FILEError(e.what());
    default: throw platanus::ErrorBase(platanus::DOUBLEHASH, e.what());
    }
}

// every entry of a host DoubleHash with a non-zero value as (key words, value); optionally zeroes the values like the
// table dump of counter.h:695-705 does
template <typename KMER, typename TABLE>
void tableEntries(const KMER &kmer, TABLE &table, bool zeroValues, std::vector<uint64_t> &words, std::vector<uint16_t> &values)
{
    char *buf = NULL;
    size_t len = 0;
    FILE *ms = open_memstream(&buf, &len);
    if (ms == NULL) throw platanus::TMPError();
    for (auto it = table.begin(), end = table.end(); it != end; ++it) {
        if (it->second == 0) continue;
        kmer.writeKey(ms, it->first);
        values.push_back(it->second);
        if (zeroValues) it->second = 0;
    }
    fclose(ms);
    words.assign(reinterpret_cast<uint64_t *>(buf), reinterpret_cast<uint64_t *>(buf + len));
    free(buf);
}

// post-condition of the counting functions: kmerFP holds one (key words, u16 count) record per distinct k-mer
// (counter.h:494-495), occurrenceDistribution = vector<u64>(65535), maxOccurrence = highest non-empty bin
inline void publish(pbk::Counter &gpu, unsigned long long kLength, FILE *kmerFP, std::vector<unsigned long long> &occurrenceDistribution,
                    unsigned long long &maxOccurrence)
{
    occurrenceDistribution = gpu.occurrenceDistribution();
    if (gpu.getNumDistinct()) maxOccurrence = gpu.getMaxOccurrence();       // left unchanged when nothing was counted (counter.h:371-376)
    gpu.exportKmers(1, /*sorted=*/false);
    const size_t W = (kLength + 31) / 32;
    const std::vector<uint64_t> &keys = gpu.keptKeys();
    const std::vector<uint16_t> &counts = gpu.keptCounts();
    for (size_t i = 0; i < counts.size(); ++i) {
        if (fwrite(&keys[i * W], sizeof(uint64_t), W, kmerFP) != W || fwrite(&counts[i], sizeof(uint16_t), 1, kmerFP) != 1)
            throw platanus::TMPError();
    }
}

// the SEQ records (common.h:426-448) of one temp file, as flat arrays
struct SeqFile {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> offsets, nposOffsets;
    std::vector<int32_t> npos;
    std::vector<long> recordStart;          // file offset of every record (+ the end), to copy kept records verbatim
    void read(FILE *fp)
    {
        offsets.assign(1, 0); nposOffsets.assign(1, 0); recordStart.clear();
        rewind(fp);
        int32_t numUnknown, length;
        for (;;) {
            recordStart.push_back(ftell(fp));
            if (fread(&numUnknown, sizeof(int32_t), 1, fp) != 1) break;
            const size_t n0 = npos.size();
            npos.resize(n0 + (size_t)numUnknown);
            if (numUnknown > 0 && fread(&npos[n0], sizeof(int32_t), (size_t)numUnknown, fp) != (size_t)numUnknown) throw platanus::ReadError();
            if (fread(&length, sizeof(int32_t), 1, fp) != 1) throw platanus::ReadError();
            const size_t b0 = bases.size();
            bases.resize(b0 + (size_t)length);
            if (length > 0 && fread(&bases[b0], 1, (size_t)length, fp) != (size_t)length) throw platanus::ReadError();
            offsets.push_back(bases.size());
            nposOffsets.push_back(npos.size());
        }
    }
};

}  // namespace shim
}  // namespace pbk

#endif  // PBK_REFERENCE_SHIM_HPP
