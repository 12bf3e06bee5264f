// TEST INFRASTRUCTURE ONLY -- drives the UNMODIFIED reference readers of seqlib.cpp (compiled from the sources where they lie under
// /root/reference; nothing is copied): ReadFastaSingleMT / ReadFastaPairMT and their tagged forms (seqlib.cpp:365-742), which the
// reference reaches from scaffold / gap_close / polish.  SURVEY.md section 8f row 3.
//   usage: ref_seqlib_harness single|pair|single_tagged|pair_tagged NUM_THREAD IS_MATE IS_FASTQ NOT_PAIR OUT_PREFIX FILE1 [FILE2]
//   -> OUT_PREFIX.<i> = the bytes of lib[i].pairFP, stdout: "numPair <n> totalLength <n>" (or "error <id>" for a platanus error)
#include "seqlib.h"

#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

int main(int argc, char **argv)
{
    if (argc < 8) return 2;
    const std::string mode = argv[1], prefix = argv[6];
    const int numThread = atoi(argv[2]);
    const bool isMate = atoi(argv[3]) != 0, isFastq = atoi(argv[4]) != 0, notPair = atoi(argv[5]) != 0;
    platanus::setGlobalTmpFileDir(".");
    std::vector<SeqLib> lib(numThread);
    for (int i = 0; i < numThread; ++i) lib[i].pairFP = platanus::makeTemporaryFile();
    try {
        std::unordered_map<std::string, int> tags;
        if (mode == "single_tagged" || mode == "pair_tagged") {
            std::vector<std::string> names;
            for (int i = 7; i < argc; ++i) names.push_back(argv[i]);
            setTagStringConverter(names, tags);
        }
        if (mode == "single") ReadFastaSingleMT(lib, argv[7], numThread, isMate, isFastq, notPair);
        else if (mode == "pair") ReadFastaPairMT(lib, argv[7], argv[8], numThread, isMate, isFastq);
        else if (mode == "single_tagged") ReadFastaSingleTaggedMT(lib, argv[7], numThread, isMate, isFastq, notPair, tags);
        else if (mode == "pair_tagged") ReadFastaPairTaggedMT(lib, argv[7], argv[8], numThread, isMate, isFastq, tags);
        else return 2;
        if (mode == "single_tagged" || mode == "pair_tagged") {      // the tag table, so that the other side can use the same ids
            FILE *tf = fopen((prefix + ".tags").c_str(), "w");
            for (auto it = tags.begin(); it != tags.end(); ++it) fprintf(tf, "%s\t%d\n", it->first.c_str(), it->second);
            fclose(tf);
        }
    } catch (platanus::ErrorBase &e) {
        std::cout << "error " << e.getID() << std::endl;
        return 0;
    }
    for (int i = 0; i < numThread; ++i) {
        FILE *out = fopen((prefix + "." + std::to_string(i)).c_str(), "wb");
        rewind(lib[i].pairFP);
        char buf[65536];
        size_t got;
        while ((got = fread(buf, 1, sizeof buf, lib[i].pairFP)) > 0) fwrite(buf, 1, got, out);
        fclose(out);
    }
    std::cout << "numPair " << lib[0].getNumPair() << " totalLength " << lib[0].getTotalLength() << std::endl;
    return 0;
}
