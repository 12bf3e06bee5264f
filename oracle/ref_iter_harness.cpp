// TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference template Counter<KMER> (counter.h, compiled from the sources
// where they lie under /root/reference; nothing is copied) through the two table-consuming steps of the iterative-k
// assembly, which the reference only reaches from inside its graph stages:
//   pickup : Counter::pickupReadMatchedEdgeKmer (counter.h:870-910)  -- keep the reads with a k-mer in the table
//   count  : Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750) -- table entries keep their
//            values, k-mers of the reads that are not in the table are counted
// The table is loaded with the reference's own Counter::readOccurrenceTableBinary from a PREFIX_kmer_occ.bin.
// Reads: one sequence per line (ACGTN), converted and dealt to NUM_THREAD SEQ temp files the way
// Assemble::readFastaUncompressed does (assemble.cpp:836-845).
//   contig : Counter::makeKmerReadDistributionFromContig (counter.h:511-593) -- table built from contigs and their coverages
//            usage: ref_iter_harness contig K CONTIGS.fa MIN_OCCURRENCE OUT   (raw kmerFP records, stdout: maxOccurrence)
//   usage: ref_iter_harness pickup|count TABLE.bin READS.txt NUM_THREAD OUT
//   pickup -> OUT: the surviving reads, one per line, per temp file in order (file 0 first)
//   count  -> OUT: the raw kmerFP records (key words + u16), then stdout: "maxOccurrence <n>"
//   flags  : the eight neighbour probes BruijnGraph::makeInitialBruijnGraph makes per k-mer of sortedKeyFP (graph.h:337-375), with
//            the reference's own KMER primitives and Counter::findValue on a table loaded by readOccurrenceTableBinary
//            usage: ref_iter_harness flags TABLE.bin - 1 OUT   (OUT: per key in ascending order, key words + one byte
//            (leftFlags << 4) | rightFlags -- the value Junction::out gets, graph.h:398)
// Built with g++ -fno-access-control (kmerLength etc. are private).
#include "counter.h"

#include <fstream>
#include <iostream>
#include <string>
#include <vector>

template <typename KMER>
static int run(const std::string &mode, const std::string &bin, const std::string &readsPath, unsigned long long numThread,
               const std::string &out)
{
    Counter<KMER> counter;
    counter.readOccurrenceTableBinary(bin);
    const unsigned k = counter.getKmerLength();

    std::vector<FILE *> readFP(numThread);
    for (unsigned long long i = 0; i < numThread; ++i) readFP[i] = platanus::makeTemporaryFile();
    {
        std::ifstream ifs(readsPath.c_str());
        std::string read;
        platanus::SEQ seq;
        unsigned long long i = 0;
        while (std::getline(ifs, read)) {
            seq.convertFromString(read);
            seq.writeTemporaryFile(readFP[i]);
            i = (i + 1) % numThread;
        }
    }

    if (mode == "pickup") {
        for (unsigned long long i = 0; i < numThread; ++i) counter.pickupReadMatchedEdgeKmer(&readFP[i]);
        std::ofstream ofs(out.c_str());
        platanus::SEQ seq;
        for (unsigned long long i = 0; i < numThread; ++i) {
            rewind(readFP[i]);
            while (seq.readTemporaryFile(readFP[i])) {
                std::string s(seq.length, 'A');
                for (long j = 0; j < seq.length; ++j) s[j] = "ACGT"[seq.base[j] & 3];
                for (int j = 0; j < seq.numUnknown; ++j) s[seq.positionUnknown[j]] = 'N';
                ofs << s << '\n';
            }
        }
        return 0;
    }
    counter.makeKmerReadDistributionConsideringPreviousGraph(k, readFP.data(), 1000000000ull, numThread);
    FILE *fp = counter.kmerFP;
    fflush(fp);
    rewind(fp);
    std::ofstream ofs(out.c_str(), std::ios::binary);
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, fp)) > 0) ofs.write(buf, got);
    std::cout << "maxOccurrence " << counter.getMaxOccurrence() << std::endl;
    return 0;
}

template <typename KMER>
static int run_flags(const std::string &bin, const std::string &out)
{
    Counter<KMER> counter;
    counter.readOccurrenceTableBinary(bin);
    const unsigned k = counter.getKmerLength();
    const unsigned long long mask = k >= 32 ? ~static_cast<unsigned long long>(0) : ~(~static_cast<unsigned long long>(0) << (2 * k));
    std::vector<typename KMER::keyType> keys;
    for (auto it = counter.occurrenceTable.begin(), end = counter.occurrenceTable.end(); it != end; ++it)
        if (it->second != 0) keys.push_back(it->first);
    std::sort(keys.begin(), keys.end());                                  // what sortedKeyFromKmerFile hands to the graph builder
    FILE *fp = fopen(out.c_str(), "wb");
    KMER startKmer(k), leftKmer(k), rightKmer(k);
    for (size_t i = 0; i < keys.size(); ++i) {
        startKmer.forward = keys[i];
        startKmer.reverseComplement();
        // graph.h:340-357
        leftKmer = startKmer;
        leftKmer.forward >>= 2;
        leftKmer.reverse <<= 2;
        leftKmer.maskReverse(mask);
        unsigned char leftFlags = 0, rightFlags = 0;
        for (unsigned char base = 0; base < 4; ++base) {
            leftKmer.setForward(leftKmer.kmerLength - 1, base);
            leftKmer.setReverse(0, 0x3 ^ base);
            if (counter.findValue(std::min(leftKmer.forward, leftKmer.reverse)) != 0) leftFlags |= 1 << base;
        }
        // graph.h:360-375
        rightKmer = startKmer;
        rightKmer.forward <<= 2;
        rightKmer.maskForward(mask);
        rightKmer.reverse >>= 2;
        for (unsigned char base = 0; base < 4; ++base) {
            rightKmer.setForward(0, base);
            rightKmer.setReverse(rightKmer.kmerLength - 1, 0x3 ^ base);
            if (counter.findValue(std::min(rightKmer.forward, rightKmer.reverse)) != 0) rightFlags |= 1 << base;
        }
        startKmer.writeKey(fp, keys[i]);
        const unsigned char o = (leftFlags << 4) | rightFlags;
        fwrite(&o, 1, 1, fp);
    }
    fclose(fp);
    return 0;
}

template <typename KMER>
static int run_contig(unsigned long long k, const std::string &fa, unsigned long long minOccurrence, const std::string &out)
{
    platanus::Contig contig;
    contig.readFastaCoverage(fa);
    Counter<KMER> counter(k);
    counter.makeKmerReadDistributionFromContig(contig, k, minOccurrence, 100000000ull);
    FILE *fp = counter.kmerFP;
    fflush(fp);
    rewind(fp);
    std::ofstream ofs(out.c_str(), std::ios::binary);
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, fp)) > 0) ofs.write(buf, got);
    std::cout << "maxOccurrence " << counter.getMaxOccurrence() << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc != 6) return 2;
    if (std::string(argv[1]) == "contig") {
        platanus::setGlobalTmpFileDir(".");
        const unsigned long long k = strtoull(argv[2], NULL, 10), minOcc = strtoull(argv[4], NULL, 10);
        const std::string fa = argv[3], out = argv[5];
        if (k <= 32) return run_contig<Kmer31>(k, fa, minOcc, out);
        if (k <= 64) return run_contig<KmerN<Binstr63> >(k, fa, minOcc, out);
        if (k <= 96) return run_contig<KmerN<Binstr95> >(k, fa, minOcc, out);
        if (k <= 128) return run_contig<KmerN<Binstr127> >(k, fa, minOcc, out);
        if (k <= 160) return run_contig<KmerN<Binstr159> >(k, fa, minOcc, out);
        return run_contig<KmerN<binstr_t> >(k, fa, minOcc, out);
    }
    const std::string mode = argv[1], bin = argv[2], reads = argv[3], out = argv[5];
    const unsigned long long numThread = strtoull(argv[4], NULL, 10);
    platanus::setGlobalTmpFileDir(".");
    omp_set_num_threads(numThread);
    const unsigned long long k = platanus::getKmerLengthFromBinary(bin);
    if (mode == "flags") {
        if (k <= 32) return run_flags<Kmer31>(bin, out);
        if (k <= 64) return run_flags<KmerN<Binstr63> >(bin, out);
        if (k <= 96) return run_flags<KmerN<Binstr95> >(bin, out);
        if (k <= 128) return run_flags<KmerN<Binstr127> >(bin, out);
        if (k <= 160) return run_flags<KmerN<Binstr159> >(bin, out);
        return run_flags<KmerN<binstr_t> >(bin, out);
    }
    if (k <= 32) return run<Kmer31>(mode, bin, reads, numThread, out);
    if (k <= 64) return run<KmerN<Binstr63> >(mode, bin, reads, numThread, out);
    if (k <= 96) return run<KmerN<Binstr95> >(mode, bin, reads, numThread, out);
    if (k <= 128) return run<KmerN<Binstr127> >(mode, bin, reads, numThread, out);
    if (k <= 160) return run<KmerN<Binstr159> >(mode, bin, reads, numThread, out);
    return run<KmerN<binstr_t> >(mode, bin, reads, numThread, out);
}
