// TEST ONLY: prints the reads pbk_ingest.hpp produces for a (plain, gzip or bzip2) file cut into T ranges, one read per line (workers run one
// after the other here, so the output is in file order), to be compared with the serial parse and with the oracle.
#include "../../platanus_b_b200/host/pbk_ingest.hpp"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

struct PrintSink : pbk::ingest::ReadSink {
    void emit_span(const char *s, size_t len) { fwrite(s, 1, len, stdout); fputc('\n', stdout); }
};

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    pbk::ingest::MappedFile f;                              // plain, gzip or bzip2
    if (f.map(argv[1], argc > 4 ? argv[4] : "/tmp") != 0) return 3;
    struct { const char *p; size_t n; const char *data() const { return p; } size_t size() const { return n; } } buf = {f.p, f.n};
    const bool fastq = std::string(argv[2]) == "fq";
    const unsigned T = (unsigned)atoi(argv[3]);
    const pbk::ingest::Plan pl = pbk::ingest::plan_ranges(buf.data(), buf.size(), fastq, T);
    for (unsigned t = 0; t < pl.n_workers; ++t) {
        PrintSink sink;
        if (pl.s[t] < pl.s[t + 1] || t == pl.final_owner)
            pbk::ingest::parse_range(buf.data(), pl.s[t], pl.s[t + 1], fastq, t == pl.final_owner, sink);
    }
    return 0;
}
