// knn_team.cu -- exact kNN, one TEAM of 16 lanes per query (two queries per warp, lockstep).
//
// The public NearestNeighborSearch path (b200vor_knn / b200vor_knn_points; reference
// KdTree::get_nearest_neighbors, points/kd_tree.cpp:156-387) on a single-level, non-periodic grid
// and k <= 40.  Everything else (two-level grids, periodic tables of the fallback cell pipeline,
// larger k) stays with the thread-per-query heap kernel of knn.cu.
//
// Per query: the rows of the (2 ring + 1)^3 block of bins around it are contiguous spans of the
// sorted seed array; the team walks them with coalesced 16-byte loads (one candidate per lane),
// computes the squared distance EXACTLY as the reference does (((dx*dx)+dy*dy)+dz*dz, individually
// rounded, no FMA: basic/geometry_nd.h:65-73 under -ffp-contract=off) and keeps the candidates with
// d2 < hi in a 64-entry pool (ballot compaction, no atomics).  hi starts at the radius of the ball
// expected to hold ~1.6 k seeds at the local density and is bisected until k <= |pool| <= 64; it
// never exceeds the distance to the unscanned region, so the pool holds EVERY seed closer than hi
// and its k smallest entries are the exact answer.  The pool is sorted by (d2 bits, user index)
// with an in-register bitonic network: ties come out by ascending index, results are
// reproducible.  Rows are written with coalesced stores.
#include "common.cuh"

namespace b200vor {

namespace {

constexpr int kG = 16;             // lanes per query
constexpr int kPool = 64;
constexpr int kWarps = 8;          // warps per block
constexpr uint32_t kFullMask = 0xffffffffu;

struct TeamParams {
    GridDesc G;
    const Seed* sorted;
    const uint32_t* cell_start;
    const uint32_t* perm;
    uint32_t n;
    KnnJob job;
    uint32_t* fail;                // number of queries this kernel gave up on (massive exact ties)
};

struct __align__(16) TeamSh {
    double pool_d2[kPool];
    uint32_t pool_j[kPool];
    uint32_t desc[kG][3];          // scan rows: prefix end, first seed, count
};

__device__ inline uint32_t team_ballot(bool p, uint32_t tbase) {
    return (__ballot_sync(kFullMask, p) >> tbase) & ((1u << kG) - 1u);
}

__device__ inline uint32_t team_exscan(uint32_t v, uint32_t lt, uint32_t& total) {
    uint32_t inc = v;
#pragma unroll
    for(int o = 1; o < kG; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFullMask, inc, o, kG);
        if(lt >= uint32_t(o)) {
            inc += t;
        }
    }
    total = __shfl_sync(kFullMask, inc, kG - 1, kG);
    return inc - v;
}

// squared distance from the query to the nearest face of the scanned block with unscanned space
// behind it (1e300: the block covers the grid)
__device__ inline double block_cover2(const GridDesc& Gd, const double q[3], const int cc[3], int r) {
    double dmin = 1e300;
#pragma unroll
    for(int a = 0; a < 3; ++a) {
        if(cc[a] - r > 0) {
            dmin = fmin(dmin, q[a] - (Gd.lo[a] + (cc[a] - r) * Gd.w[a]));
        }
        if(cc[a] + r < Gd.g[a] - 1) {
            dmin = fmin(dmin, (Gd.lo[a] + (cc[a] + r + 1) * Gd.w[a]) - q[a]);
        }
    }
    if(dmin >= 1e300) {
        return 1e300;
    }
    dmin -= 1e-9 * Gd.w[0];
    return dmin > 0.0 ? dmin * dmin : 0.0;
}

struct Key {
    unsigned long long d;   // bits of the (non-negative) squared distance: ordered as the doubles are
    uint32_t id;            // user index
};

__device__ inline bool key_less(const Key& a, const Key& b) {
    return a.d < b.d || (a.d == b.d && a.id < b.id);
}

// in-register bitonic sort of kPool keys, element e = r * kG + lane (ascending in e)
__device__ inline void sort_pool(uint32_t lt, Key (&key)[kPool / kG]) {
    constexpr int RK = kPool / kG;
#pragma unroll
    for(int k = 2; k <= kPool; k <<= 1) {
#pragma unroll
        for(int j = k >> 1; j > 0; j >>= 1) {
            if(j >= kG) {
                const int jr = j / kG;
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    if((r & jr) == 0) {
                        const bool asc = (((r * kG) & k) == 0);
                        const bool sw = key_less(key[r | jr], key[r]) == asc;
                        if(sw) {
                            const Key t = key[r];
                            key[r] = key[r | jr];
                            key[r | jr] = t;
                        }
                    }
                }
            } else {
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    Key o;
                    o.d = __shfl_xor_sync(kFullMask, key[r].d, j, kG);
                    o.id = __shfl_xor_sync(kFullMask, key[r].id, j, kG);
                    const uint32_t e = uint32_t(r * kG) + lt;
                    const bool keep_min = (((lt & uint32_t(j)) == 0) == ((e & uint32_t(k)) == 0));
                    const bool o_less = key_less(o, key[r]);
                    if(keep_min == o_less) {
                        key[r] = o;
                    }
                }
            }
        }
    }
}

__device__ inline void sort_pool32(uint32_t lt, uint32_t (&key)[kPool / kG]) {
    constexpr int RK = kPool / kG;
#pragma unroll
    for(int k = 2; k <= kPool; k <<= 1) {
#pragma unroll
        for(int j = k >> 1; j > 0; j >>= 1) {
            if(j >= kG) {
                const int jr = j / kG;
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    if((r & jr) == 0) {
                        const bool asc = (((r * kG) & k) == 0);
                        const uint32_t lo = min(key[r], key[r | jr]), hi = max(key[r], key[r | jr]);
                        key[r] = asc ? lo : hi;
                        key[r | jr] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    const uint32_t other = __shfl_xor_sync(kFullMask, key[r], j, kG);
                    const uint32_t e = uint32_t(r * kG) + lt;
                    const bool keep_min = (((lt & uint32_t(j)) == 0) == ((e & uint32_t(k)) == 0));
                    key[r] = keep_min ? min(key[r], other) : max(key[r], other);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(32 * kWarps) knn_team_kernel(TeamParams P) {
    constexpr int TPW = 32 / kG;
    constexpr int RK = kPool / kG;
    __shared__ TeamSh shm[kWarps * TPW];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t tiw = lane / kG, lt = lane % kG, tbase = tiw * kG;
    TeamSh* sh = &shm[wid * TPW + tiw];
    const GridDesc& Gd = P.G;
    const uint32_t t = (blockIdx.x * kWarps + wid) * TPW + tiw;
    const bool alive = t < P.job.nq;
    const uint32_t tq = alive ? t : 0u;   // a team without a query repeats query 0 and writes nothing
    const uint32_t k = P.job.k;

    double q[3];
    uint32_t row;
    if(P.job.qpts) {
        q[0] = P.job.qpts[3 * size_t(tq)];
        q[1] = P.job.qpts[3 * size_t(tq) + 1];
        q[2] = P.job.qpts[3 * size_t(tq) + 2];
        row = tq;
    } else {
        const uint32_t idx = P.job.qlist ? P.job.qlist[tq] : tq;
        const Seed s = load_seed(P.sorted + idx);
        q[0] = s.x; q[1] = s.y; q[2] = s.z;
        row = __ldg(P.perm + idx);
    }
    int cc[3];
    cell_coords(Gd, q[0], q[1], q[2], cc[0], cc[1], cc[2]);
    const int rmax = max(max(cc[0], Gd.g[0] - 1 - cc[0]), max(max(cc[1], Gd.g[1] - 1 - cc[1]), max(cc[2], Gd.g[2] - 1 - cc[2])));
    int ring = min(1, rmax);
    // no seed is farther than this: stands in for "infinity" when the block covers the grid
    double far2 = 0.0;
#pragma unroll
    for(int a = 0; a < 3; ++a) {
        const double e = fmax(fabs(q[a] - Gd.lo[a]), fabs(Gd.lo[a] + Gd.g[a] * Gd.w[a] - q[a])) * 1.001;
        far2 += e * e;
    }
    far2 = fmax(far2, 1e-300);
    auto cover = [&](int r) { return fmin(block_cover2(Gd, q, cc, r), far2); };
    double cov2 = cover(ring);
    // first radius: the ball expected to hold `target` seeds at the mean density of the grid, never
    // beyond what the block covers (a wrong guess only costs a bisection step)
    const float target = fminf(1.6f * float(k) + 4.0f, 0.8f * float(kPool));
    double hi;
    {
        const float binvol = float(Gd.w[0] * Gd.w[1] * Gd.w[2]);
        const float rr = cbrtf(target * 0.2387324146378430f * binvol * float(Gd.ncell) / float(max(P.n, 1u)));
        hi = fmin(fmax(double(rr * rr), 1e-300), cov2);
    }
    double hlo = 0.0, hhi = 1e300;       // a radius known to hold < k seeds / more than the pool (1e300: none yet)
    bool done = false, failed = false;
    uint32_t count = 0;
    int guard = 0;
    while(__any_sync(kFullMask, !done)) {
        if(!done) {
            count = 0;   // (a finished team keeps its pool while the other team of the warp goes on)
        }
        const int R1 = 2 * ring + 1, nrows = R1 * R1;
        for(int rb = 0; __any_sync(kFullMask, !done && rb < nrows); rb += kG) {
            const bool rowact = !done && rb < nrows;
            const int rowi = rb + int(lt);
            uint32_t a0 = 0, n0 = 0;
            if(rowact && rowi < nrows) {
                const int dy = rowi % R1 - ring, dz = rowi / R1 - ring;
                const int y = cc[1] + dy, z = cc[2] + dz;
                if(y >= 0 && y < Gd.g[1] && z >= 0 && z < Gd.g[2]) {
                    // the row of bins is scanned only where it meets the ball d2 < hi
                    const double slack = 1e-9 * Gd.w[0];
                    double dyy = dy == 0 ? 0.0 : (dy < 0 ? q[1] - (Gd.lo[1] + (y + 1) * Gd.w[1]) : (Gd.lo[1] + y * Gd.w[1]) - q[1]);
                    double dzz = dz == 0 ? 0.0 : (dz < 0 ? q[2] - (Gd.lo[2] + (z + 1) * Gd.w[2]) : (Gd.lo[2] + z * Gd.w[2]) - q[2]);
                    dyy = fmax(dyy - slack, 0.0);
                    dzz = fmax(dzz - slack, 0.0);
                    const double rem = hi - (dyy * dyy + dzz * dzz);
                    if(rem > 0.0) {
                        const double rx = sqrt(rem) + slack;
                        const int xa = max(max(cc[0] - ring, 0), int(floor((q[0] - rx - Gd.lo[0]) * Gd.inv_w[0])));
                        const int xb = min(min(cc[0] + ring, Gd.g[0] - 1), int(floor((q[0] + rx - Gd.lo[0]) * Gd.inv_w[0])));
                        if(xa <= xb) {
                            const uint32_t base = uint32_t(Gd.g[0]) * (uint32_t(y) + uint32_t(Gd.g[1]) * uint32_t(z));
                            a0 = __ldg(P.cell_start + base + xa);
                            n0 = __ldg(P.cell_start + base + xb + 1) - a0;
                        }
                    }
                }
            }
            uint32_t total;
            const uint32_t pre = team_exscan(n0, lt, total);
            sh->desc[lt][0] = pre + n0;
            sh->desc[lt][1] = a0;
            sh->desc[lt][2] = n0;
            __syncwarp();
            // ---- candidates of the rows -> pool ----
            uint32_t cur = 0, cur_end = sh->desc[0][0], cur_beg = 0;
            for(uint32_t c0 = 0; __any_sync(kFullMask, rowact && c0 < total); c0 += kG) {
                const uint32_t c = c0 + lt;
                bool in = false;
                uint32_t j = 0;
                double d2 = 0.0;
                if(rowact && c < total) {
                    while(c >= cur_end) {
                        cur_beg = cur_end;
                        ++cur;
                        cur_end = sh->desc[cur][0];
                    }
                    j = sh->desc[cur][1] + (c - cur_beg);
                    const Seed s = load_seed(P.sorted + j);
                    const double dx = __dsub_rn(s.x, q[0]), dy = __dsub_rn(s.y, q[1]), dz = __dsub_rn(s.z, q[2]);
                    d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    in = d2 < hi || hi >= far2;
                }
                const uint32_t bal = team_ballot(in, tbase);
                if(in) {
                    const uint32_t pos = count + __popc(bal & ((1u << lt) - 1u));
                    if(pos < uint32_t(kPool)) {
                        sh->pool_d2[pos] = d2;
                        sh->pool_j[pos] = j;
                    }
                }
                count += __popc(bal);
            }
            __syncwarp();
        }
        if(!done) {
            if(count > uint32_t(kPool)) {
                hhi = hi;
                hi = 0.5 * (hlo + hhi);
            } else if(count >= k) {
                done = true;                     // hi <= cov2: every seed closer than hi is in the pool
            } else if(cov2 >= far2 && hi >= cov2) {
                done = true;                     // the whole grid was scanned (k > n cannot happen: checked by the caller)
            } else {
                hlo = hi;
                if(hhi < 1e300) {
                    hi = 0.5 * (hlo + hhi);
                } else if(hi < cov2) {
                    hi = fmin(fmax(1.7 * hi, 1e-300), cov2);
                } else {
                    ring = min(ring + 1, max(rmax, 1));
                    cov2 = cover(ring);
                    hi = fmin(fmax(1.7 * hi, 1e-300), cov2);
                }
            }
            if(++guard > 200 && !done) {
                failed = true;                   // more than a pool of seeds at one distance: heap kernel
                done = true;
            }
        }
    }
    // ---- sort the pool by (d2, user index), write the first k ----
    // fast path: 32-bit keys (d2 as float, rounded down, 6 low bits = pool slot); the order is then
    // verified with the exact (d2 bits, user index) keys of neighbouring ranks, and the rare pool
    // with two candidates inside one float step is sorted again with the exact keys
    const uint32_t cnt = min(count, uint32_t(kPool));
    uint32_t k32[RK];
#pragma unroll
    for(int r = 0; r < RK; ++r) {
        const uint32_t e = lt + kG * r;
        k32[r] = 0xffffffffu;
        if(e < cnt) {
            k32[r] = (__float_as_uint(__double2float_rd(sh->pool_d2[e])) & ~63u) | e;
        }
    }
    sort_pool32(lt, k32);
    Key key[RK];
    bool bad = false;
#pragma unroll
    for(int r = 0; r < RK; ++r) {
        const uint32_t e = lt + kG * r;
        key[r].d = ~0ull;
        key[r].id = 0xffffffffu;
        if(e < cnt) {
            const uint32_t slot = k32[r] & 63u;
            key[r].d = (unsigned long long)__double_as_longlong(sh->pool_d2[slot]);
            key[r].id = __ldg(P.perm + sh->pool_j[slot]);
        }
    }
#pragma unroll
    for(int r = 0; r < RK; ++r) {
        // successor of element e = r kG + lt: lane lt + 1 of the same round, or lane 0 of the next
        Key nx;
        nx.d = __shfl_down_sync(kFullMask, key[r].d, 1, kG);
        nx.id = __shfl_down_sync(kFullMask, key[r].id, 1, kG);
        Key first_next;
        first_next.d = __shfl_sync(kFullMask, key[(r + 1) % RK].d, 0, kG);
        first_next.id = __shfl_sync(kFullMask, key[(r + 1) % RK].id, 0, kG);
        if(lt == kG - 1) {
            nx = first_next;
        }
        const uint32_t e = lt + kG * r;
        if(e + 1 < cnt && key_less(nx, key[r])) {
            bad = true;
        }
    }
    if(__any_sync(kFullMask, bad)) {
        sort_pool(lt, key);
    }
    if(alive && !failed) {
#pragma unroll
        for(int r = 0; r < RK; ++r) {
            const uint32_t e = lt + kG * r;
            if(e < k) {
                const bool have = e < count;
                P.job.out_nbr[size_t(row) * k + e] = have ? key[r].id : 0xFFFFFFFFu;
                if(P.job.out_d2) {
                    P.job.out_d2[size_t(row) * k + e] =
                        have ? __longlong_as_double((long long)key[r].d) : __longlong_as_double(0x7FF0000000000000LL);
                }
            }
        }
    }
    if(alive && failed && lt == 0) {
        atomicAdd(P.fail, 1u);
    }
    if(alive && P.job.out_flag && lt == 0) {
        P.job.out_flag[row] = failed ? 1 : 0;
    }
}

}  // namespace

bool knn_team_applicable(const Ctx* c, const KnnJob& job) {
    return !c->grid.two_level && !c->grid.periodic && job.user_ids && job.k <= 40 && job.k <= c->n;
}

// returns B200VOR_OK and *nfail = number of queries left to the heap kernel (the caller re-runs the job)
int run_knn_team(Ctx* c, const KnnJob& job, uint32_t* nfail) {
    TeamParams P;
    P.G = c->grid;
    P.sorted = c->sorted.p;
    P.cell_start = c->cell_start.p;
    P.perm = c->perm.p;
    P.n = c->n;
    P.job = job;
    int rc = c->counter.reserve(4);
    if(rc) {
        return rc;
    }
    P.fail = c->counter.p;
    B200_CUDA_CHECK(cudaMemsetAsync(c->counter.p, 0, sizeof(uint32_t), c->stream));
    const uint32_t per_block = kWarps * (32 / kG);
    const uint32_t nblk = (job.nq + per_block - 1) / per_block;
    knn_team_kernel<<<nblk, 32 * kWarps, 0, c->stream>>>(P);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(nfail, c->counter.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

}  // namespace b200vor
