// dump_mesh.cpp — prints the triangles TriangleMesh(file, a, b, ..., typeofdata) loads, one "%.17g x9" line per triangle.
// Used by tests/test_host_cpp.py to pin the C++ loader against the reference's loader arithmetic (objects.h:343-400).
#include <cstdlib>

#include "cgrt_host.hpp"

int main(int argc, char **argv) {
    if (argc != 7) { std::fprintf(stderr, "usage: %s <file> <typeofdata> <a> <bx> <by> <bz>\n", argv[0]); return 2; }
    try {
        cgrt_host::TriangleMesh m(argv[1], std::atof(argv[3]), cgrt_host::Vec3(std::atof(argv[4]), std::atof(argv[5]), std::atof(argv[6])),
                                  cgrt_host::Vec3(1, 1, 1), 0, 0, std::atoi(argv[2]));
        for (size_t i = 0; i < m.size(); i++) {
            for (int k = 0; k < 9; k++) std::printf("%.17g%c", m.tri9[9 * i + k], k == 8 ? '\n' : ' ');
        }
    } catch (const cgrt_host::Error &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
