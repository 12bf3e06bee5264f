// TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference classes (compiled from the sources where they lie under
// /root/reference; nothing is copied) to print what ContigDivider::getOccurrenceArray (kmer_divide.cpp:151-197)
// computes: the occurrence of every k-mer window of a contig file in a PREFIX_kmer_occ.bin.  The reference's own
// `kmer_divide` command computes this array and immediately folds it into break points; its dump routine
// (dumpKmerCoverage, kmer_divide.cpp:374-388) exists but the call is commented out (kmer_divide.cpp:118).  This harness
// repeats the first half of ContigDivider::exec (kmer_divide.cpp:72-116) and then calls that dump routine.
// Output format (the reference's): ">name" line per contig, then one occurrence per window start.
//   usage: ref_occ_harness KMER_OCC.bin CONTIGS.fa OUT.txt
// Built with g++ -fno-access-control: the members ContigDivider::exec works on are private.
#include "kmer_divide.cpp"      // -I/root/reference: the reference source itself, not a copy

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    const std::string bin = argv[1], fa = argv[2], out = argv[3];
    ContigDivider d;
    omp_set_num_threads(1);
    const unsigned long long kmerLength = platanus::getKmerLengthFromBinary(bin);
    d.contig.readFastaCoverage(fa);
    if (kmerLength <= 32) { Counter<Kmer31> c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    else if (kmerLength <= 64) { Counter<KmerN<Binstr63> > c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    else if (kmerLength <= 96) { Counter<KmerN<Binstr95> > c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    else if (kmerLength <= 128) { Counter<KmerN<Binstr127> > c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    else if (kmerLength <= 160) { Counter<KmerN<Binstr159> > c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    else { Counter<KmerN<binstr_t> > c(kmerLength); c.readOccurrenceTableBinary(bin); d.getOccurrenceArray(c); }
    d.dumpKmerCoverage(out);
    return 0;
}
