// grid_types.h -- plain C++ types shared by the CUDA kernels and by the host-side emulation of the
// thread-per-cell kernel (tests/host/tpc_host.cpp): no CUDA-only construct in here.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__
#else
#define B200_HD
#endif

namespace b200vor {

// Uniform grid over the domain.  Cell id = ix + gx*(iy + gy*iz): x is the fastest
// axis, so a run of cells along x is one contiguous span of the sorted point array.
struct GridDesc {
    double lo[3];
    double w[3];       // cell width
    double inv_w[3];
    int g[3];
    int periodic;
    double period;
    uint32_t ncell;
    // two-level grid (clustered seeds): a bin with more than kDenseBin seeds carries its own
    // s x s x s sub-grid, s = 1 << ls; sub[b] = (first virtual cell of bin b) << 3 | ls, and
    // vstart[] is the start table of the virtual cells (x fastest inside a bin)
    int two_level;
    uint32_t nvirt;
    // slab-local grid (slab.cu): bit f of `open` = face f of the grid (x-,x+,y-,y+,z-,z+) is not a
    // face of the domain: seeds this rank does not hold lie behind it, so the scanned block never
    // covers more than the grid there and a cell that still needs more is left UNCERTIFIED
    int open;
};
constexpr uint32_t kDenseBin = 32;   // a uniform (Poisson, 8 per bin) input never reaches it
constexpr uint32_t kSubPpc = 8;      // sub-grid sized for <= kSubPpc seeds per virtual cell
constexpr int kMaxSubLog = 5;        // s <= 32
constexpr int kSmallPeriodicGrid = 12; // periodic grids with fewer bins per axis take the two-level walk
constexpr int kSuperLog = 3;           // two-level grid: occupancy of 8 x 8 x 8 blocks of bins (far-reaching cells skip empty space)
constexpr int kWlocRingNear = 3;     // bins of dilation of the two local weight bounds (Laguerre)
constexpr int kWlocRingFar = 8;

// One seed in grid order: 32 bytes, read with two 16-byte loads.
struct alignas(16) Seed {
    double x, y, z, w;
};

B200_HD inline int clampi(int v, int lo, int hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}

// delaunay/periodic.h:104-113 + periodic.cpp:80-84: instance of a translation
B200_HD inline uint32_t translation_to_instance(int tx, int ty, int tz) {
    int i = (tz + 1) + 3 * (ty + 1) + 9 * (tx + 1);
    // reorder_instances: 13 -> 0, 0..12 -> 1..13, 14..26 unchanged
    return uint32_t(i == 13 ? 0 : (i < 13 ? i + 1 : i));
}


}  // namespace b200vor
