// Mutation fuzzer for the host side of jsrt_scene_create (wire reader, flattener, SDF compiler, world boxes): a malformed
// blob must end in an error message, never in a crash of the host process (Node, in the reference's deployment).
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -Ijsraytracer_b200/csrc -o /tmp/fuzz_wire tools/fuzz_wire.cpp \
//       jsraytracer_b200/csrc/{wire,scene_flatten,sdf_compile,bvh_build}.cpp
//   /tmp/fuzz_wire <base blob> <format 0|1> <seed> <cases>
// Found in round 1: unbounded recursion through cyclic `_r` references in the kdtree / SDF tree, a msgpack array count
// reserved before it was checked against the remaining input, unbounded SDF unrolling (tests/test_wire_robustness.py).
#include <cstdio>
#include <fstream>
#include <iterator>
#include <random>
#include <vector>
#include "host_scene.h"
using namespace jsrt;
int main(int argc, char** argv) {
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<unsigned char> base((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    int fmt = atoi(argv[2]); unsigned seed = atoi(argv[3]); int n = atoi(argv[4]);
    std::mt19937 rng(seed);
    int ok = 0, bad = 0;
    for (int it = 0; it < n; ++it) {
        std::vector<unsigned char> b = base;
        int k = 1 << (rng() % 5);
        int mode = rng() % 100;
        for (int j = 0; j < k && !b.empty(); ++j) {
            size_t pos = rng() % b.size();
            if (mode < 45) b[pos] = (unsigned char)rng();
            else if (mode < 60) { size_t len = 1 + rng() % 64; b.erase(b.begin() + pos, b.begin() + std::min(b.size(), pos + len)); }
            else if (mode < 75) { for (int q = 1 + rng() % 8; q > 0; --q) b.insert(b.begin() + pos, (unsigned char)rng()); }
            else if (mode < 85) { b.resize(pos); }
            else { // copy a chunk over another place (structure-preserving-ish)
                size_t src = rng() % b.size(), len = 1 + rng() % 256;
                for (size_t q = 0; q < len && src + q < b.size() && pos + q < b.size(); ++q) b[pos + q] = b[src + q];
            }
        }
        if (b.empty()) b.push_back(0);
        if (argc > 5) { std::ofstream o(argv[5], std::ios::binary); o.write((const char*)b.data(), b.size()); }      // keep the current case for a post-mortem
        try {
            WireDoc doc(b.data(), b.size(), fmt);
            HostScene hs; flattenScene(doc, hs);
            std::vector<float> wb; computeWorldBoxes(hs, wb);
            ++ok;
        } catch (const std::exception& e) { ++bad; }
    }
    printf("%s fmt %d seed %u: accepted %d rejected %d\n", argv[1], fmt, seed, ok, bad);
    return 0;
}
