// Times (and, built with -pg, profiles) the host side of jsrt_scene_create — wire reader + flattener — on a blob file:
//   g++ -O2 -std=c++17 -Ijsraytracer_b200/csrc -o /tmp/host_create_prof tools/host_create_prof.cpp \
//       jsraytracer_b200/csrc/{wire,scene_flatten,sdf_compile,bvh_build,obj_parse}.cpp
//   /tmp/host_create_prof <blob> <format 0 = JSON | 1 = msgpack>
// The numbers of INTEGRATION.md §4a come from the library itself (JSRT_HOST_TIMING=1); this driver is what the gprof
// profile behind this round's reader changes was taken with.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>
#include "wire.h"
#include "host_scene.h"
using namespace jsrt;
int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <blob> <format>\n", argv[0]); return 2; }
    std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
    std::vector<char> b((size_t)f.tellg());
    f.seekg(0);
    f.read(b.data(), (std::streamsize)b.size());
    int fmt = atoi(argv[2]);
    for (int it = 0; it < 3; ++it) {
        auto t0 = std::chrono::steady_clock::now();
        WireDoc doc((const uint8_t*)b.data(), b.size(), fmt);
        auto t1 = std::chrono::steady_clock::now();
        HostScene hs; flattenScene(doc, hs);
        auto t2 = std::chrono::steady_clock::now();
        printf("parse %.3f flatten %.3f\n", std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count());
    }
}
