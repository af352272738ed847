// CPU-only check of vpb200::Notes (vp_facade.hpp, the reference's Notes interface: prepare / getClosestFreq,
// Source/Notes.h:27-28): prints the table of a key and the closest frequency for a sweep of pitches, one value per line
// with 17 significant digits (tests/test_host.py compares them with the oracle and the reference build).
#include <cstdio>
#include <cstdlib>

#include "vp_facade.hpp"

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: notes_check key fMin fMax [pitch ...]\n"); return 2; }
    const int k = std::atoi(argv[1]);
    vpb200::Notes n;
    n.prepare((vpb200::key)k, std::atof(argv[2]), std::atof(argv[3]));
    std::printf("%zu\n", n.size());
    for (size_t i = 0; i <= n.size(); ++i) std::printf("%.17g\n", n.data()[i]);  // incl. the popped slot
    for (int i = 4; i < argc; ++i) {
        const double p = std::atof(argv[i]);
        std::printf("%.17g\n", n.getClosestFreq(p, (vpb200::key)k));
    }
    // key change through getClosestFreq rebuilds the table (Notes.cpp:83-88)
    const int k2 = (k + 5) % 13;
    std::printf("%.17g\n", n.getClosestFreq(220.0, (vpb200::key)k2));
    std::printf("%zu\n", n.size());
    return 0;
}
