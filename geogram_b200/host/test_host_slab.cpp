// test_host_slab.cpp -- multi-GPU Lloyd relaxation driven from plain C++ (no torch, no MPI): one
// process per GPU, started by hand or by tests/test_host_cpp.py as
//     test_host_slab <rank> <world> <id-file> [n_total] [iterations]
// Rank 0 creates the NCCL id (b200vor_comm_unique_id) and writes it to <id-file>; the other ranks
// wait for the file.  Every rank generates the same seeds, sorts them by x, owns one slab
// (b200vor_slab_set_points), runs b200vor_slab_lloyd, and checks its rows against a single-context
// run over ALL seeds on its own GPU (b200vor_lloyd): the multi-GPU form of
// CentroidalVoronoiTesselation::Lloyd_iterations (voronoi/CVT.cpp:133-169).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../include/b200vor.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc__ = (call);                                                               \
        if(rc__ != B200VOR_OK) {                                                         \
            std::fprintf(stderr, "%s -> %d: %s\n", #call, rc__, b200vor_last_error());   \
            return 1;                                                                    \
        }                                                                                \
    } while(0)

int main(int argc, char** argv) {
    if(argc < 4) {
        std::fprintf(stderr, "usage: test_host_slab <rank> <world> <id-file> [n_total] [iterations]\n");
        return 2;
    }
    const int rank = std::atoi(argv[1]), world = std::atoi(argv[2]);
    const char* idfile = argv[3];
    const uint32_t n = argc > 4 ? uint32_t(std::atoi(argv[4])) : 300000u;
    const uint32_t iters = argc > 5 ? uint32_t(std::atoi(argv[5])) : 3u;

    // the same seeds on every rank, sorted by x: rank r owns the r-th slab of equal count
    std::mt19937_64 rng(2024);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double> xyz(3 * size_t(n));
    for(double& v : xyz) {
        v = U(rng);
    }
    std::vector<uint32_t> order(n);
    for(uint32_t i = 0; i < n; ++i) {
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return xyz[3 * size_t(a)] < xyz[3 * size_t(b)]; });
    std::vector<double> sorted(3 * size_t(n));
    for(uint32_t i = 0; i < n; ++i) {
        std::memcpy(&sorted[3 * size_t(i)], &xyz[3 * size_t(order[i])], 3 * sizeof(double));
    }
    const uint32_t base = n / uint32_t(world), rem = n % uint32_t(world);
    const uint32_t first = uint32_t(rank) * base + std::min(uint32_t(rank), rem);
    const uint32_t n_own = base + (uint32_t(rank) < rem ? 1u : 0u);

    // the NCCL id: 128 bytes from rank 0 to everybody, here through a file
    unsigned char id[B200VOR_UNIQUE_ID_BYTES];
    if(rank == 0) {
        CHECK(b200vor_comm_unique_id(id));
        std::string tmp = std::string(idfile) + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        std::fwrite(id, 1, sizeof(id), f);
        std::fclose(f);
        std::rename(tmp.c_str(), idfile);
    } else {
        FILE* f = nullptr;
        for(int t = 0; t < 600 && !(f = std::fopen(idfile, "rb")); ++t) {
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
        if(!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) {
            std::fprintf(stderr, "rank %d: no NCCL id in %s\n", rank, idfile);
            return 1;
        }
        std::fclose(f);
    }

    b200vor_ctx* slab = nullptr;
    CHECK(b200vor_create(&slab, rank));          // one GPU per rank
    CHECK(b200vor_set_domain_periodic(slab, 1.0));
    CHECK(b200vor_comm_init(slab, id, rank, world));
    CHECK(b200vor_slab_set_points(slab, n_own, first, n, &sorted[3 * size_t(first)], nullptr, 0));
    std::vector<double> mine(3 * size_t(n_own));
    CHECK(b200vor_slab_lloyd(slab, iters, nullptr, mine.data(), 0));
    b200vor_slab_stats st;
    CHECK(b200vor_slab_get_stats(slab, &st));

    // the single-context answer, all seeds on this rank's GPU
    b200vor_ctx* one = nullptr;
    CHECK(b200vor_create(&one, rank));
    CHECK(b200vor_set_domain_periodic(one, 1.0));
    CHECK(b200vor_set_points(one, n, sorted.data(), 3, nullptr, 0));
    std::vector<double> ref(3 * size_t(n));
    b200vor_cell_params p = {0, 0, 0, 3, 0};
    CHECK(b200vor_lloyd(one, iters, &p, ref.data(), 0));
    double err = 0.0;
    for(size_t k = 0; k < mine.size(); ++k) {
        err = std::max(err, std::fabs(mine[k] - ref[3 * size_t(first) + k]));
    }
    std::printf("rank %d of %d: %u own seeds, %u x-ghosts, %u y/z images, halo %.4f, %llu bytes sent, max |x - x_single| = %.3e\n",
                rank, world, st.n_own, st.n_ghost_x, st.n_ghost_yz, st.halo, (unsigned long long)st.bytes_sent, err);
    b200vor_destroy(one);
    b200vor_destroy(slab);
    if(!(err < 1e-12) || (world > 1 && st.bytes_sent == 0)) {
        std::printf("host slab test: FAILED\n");
        return 1;
    }
    std::printf("host slab test: ok\n");
    return 0;
}
