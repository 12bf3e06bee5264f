// TEST DRIVER for platanus_b_b200/host/pbk_seqlib.hpp: same command line and outputs as oracle/ref_seqlib_harness.cpp (the reference's
// own readers), so that the two can be compared byte for byte.
//   seqlib_check single|pair|single_tagged|pair_tagged NUM_THREAD IS_MATE IS_FASTQ NOT_PAIR OUT_PREFIX FILE1 [FILE2]   (tagged: reads OUT_PREFIX.tags)
#include "../../platanus_b_b200/host/pbk_seqlib.hpp"

#include <fstream>
#include <iostream>

int main(int argc, char **argv)
{
    if (argc < 8) return 2;
    const std::string mode = argv[1], prefix = argv[6];
    const int numThread = atoi(argv[2]);
    const bool isMate = atoi(argv[3]) != 0, isFastq = atoi(argv[4]) != 0, notPair = atoi(argv[5]) != 0;
    pbk::seqlib::PairLibrary lib;
    for (int i = 0; i < numThread; ++i) lib.pairFP.push_back(pbk::Counter::makeTemporaryFile("."));
    try {
        std::unordered_map<std::string, int> tags;
        const bool tagged = mode == "single_tagged" || mode == "pair_tagged";
        if (tagged) {                                            // the table setTagStringConverter built on the reference side
            std::ifstream tf((prefix + ".tags").c_str());
            std::string t; int id;
            while (tf >> t >> id) tags[t] = id;
        }
        if (mode == "single" || mode == "single_tagged") pbk::seqlib::ReadFastaSingleMT(lib, argv[7], numThread, isMate, isFastq, notPair, tagged ? &tags : NULL);
        else pbk::seqlib::ReadFastaPairMT(lib, argv[7], argv[8], numThread, isMate, isFastq, tagged ? &tags : NULL);
    } catch (pbk::ErrorBase &e) {
        std::cout << "error " << e.getID() << std::endl;
        return 0;
    }
    for (int i = 0; i < numThread; ++i) {
        FILE *out = fopen((prefix + "." + std::to_string(i)).c_str(), "wb");
        rewind(lib.pairFP[i]);
        char buf[65536];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, lib.pairFP[i])) > 0) fwrite(buf, 1, got, out);
        fclose(out);
    }
    std::cout << "numPair " << lib.numPair << " totalLength " << lib.totalLength << std::endl;
    return 0;
}
