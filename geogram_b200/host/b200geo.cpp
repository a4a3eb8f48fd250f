// b200geo.cpp -- see b200geo.h.  Every device operation is one call of the C ABI
// (include/b200vor.h); nothing here computes geometry on the CPU.
#include "b200geo.h"

#include <stdio.h>
#include <string.h>
#include <math.h>

#include <fstream>

#include <algorithm>
#include <stdexcept>

#include "../../include/b200vor.h"

namespace B200GEO {

namespace {

void check(int rc, const char* what) {
    if(rc != B200VOR_OK) {
        throw std::runtime_error(std::string(what) + ": " + b200vor_last_error());
    }
}

b200vor_ctx* new_context() {
    b200vor_ctx* ctx = nullptr;
    check(b200vor_create(&ctx, 0), "b200vor_create");
    return ctx;
}

}  // namespace

// ------------------------------------------------------------------ NearestNeighborSearch
NearestNeighborSearch::NearestNeighborSearch(coord_index_t dimension)
    : dimension_(dimension), nb_points_(0), stride_(0), points_(nullptr), exact_(true),
      ctx_(new_context()), table_k_(0) {}

NearestNeighborSearch::~NearestNeighborSearch() { b200vor_destroy(ctx_); }

NearestNeighborSearch* NearestNeighborSearch::create(coord_index_t dimension, const std::string& name) {
    if(dimension != 3) {
        fprintf(stderr, "[B200NN] warning: only dimension 3 is implemented (asked %d)\n", int(dimension));
        return nullptr;
    }
    if(name != "default" && name != "B200NN") {
        // the reference falls back to its default with a logger warning (nn_search.cpp:142-149)
        fprintf(stderr, "[B200NN] warning: could not create NN search algorithm: %s, falling back to B200NN\n",
                name.c_str());
    }
    return new NearestNeighborSearch(dimension);
}

void NearestNeighborSearch::set_points(index_t nb_points, const double* points) {
    set_points(nb_points, points, dimension_);
}

void NearestNeighborSearch::set_points(index_t nb_points, const double* points, index_t stride) {
    std::lock_guard<std::mutex> g(lock_);
    nb_points_ = nb_points;
    points_ = points;
    stride_ = stride;
    table_k_ = 0;
    check(b200vor_set_points(ctx_, nb_points, points, stride, nullptr, 0), "NearestNeighborSearch::set_points");
}

void NearestNeighborSearch::get_nearest_neighbors(index_t nb_neighbors, const double* query_point,
                                                  index_t* neighbors, double* neighbors_sq_dist) const {
    std::lock_guard<std::mutex> g(lock_);
    check(b200vor_knn_points(ctx_, 1, query_point, nb_neighbors, neighbors, neighbors_sq_dist, 0),
          "NearestNeighborSearch::get_nearest_neighbors");
}

void NearestNeighborSearch::get_nearest_neighbors(index_t nb_neighbors, const double* query_point,
                                                  index_t* neighbors, double* neighbors_sq_dist,
                                                  KeepInitialValues) const {
    // the caller's (sorted) candidates stay in the running: k best of the union
    std::vector<index_t> nn(nb_neighbors);
    std::vector<double> dd(nb_neighbors);
    get_nearest_neighbors(nb_neighbors, query_point, nn.data(), dd.data());
    std::vector<std::pair<double, index_t> > all;
    for(index_t i = 0; i < nb_neighbors; ++i) {
        all.push_back(std::make_pair(dd[i], nn[i]));
        if(neighbors[i] != index_t(-1) && std::find(nn.begin(), nn.end(), neighbors[i]) == nn.end()) {
            all.push_back(std::make_pair(neighbors_sq_dist[i], neighbors[i]));
        }
    }
    std::sort(all.begin(), all.end());
    for(index_t i = 0; i < nb_neighbors; ++i) {
        neighbors[i] = all[i].second;
        neighbors_sq_dist[i] = all[i].first;
    }
}

void NearestNeighborSearch::get_nearest_neighbors(index_t nb_neighbors, index_t query_point,
                                                  index_t* neighbors, double* neighbors_sq_dist) const {
    if(query_point >= nb_points_) {
        throw std::runtime_error("NearestNeighborSearch::get_nearest_neighbors: query index out of range");
    }
    std::lock_guard<std::mutex> g(lock_);
    if(table_k_ < nb_neighbors) {
        // one launch answers this query for every stored point
        table_nbr_.resize(size_t(nb_points_) * nb_neighbors);
        table_d2_.resize(size_t(nb_points_) * nb_neighbors);
        check(b200vor_knn(ctx_, nb_neighbors, table_nbr_.data(), table_d2_.data(), 0),
              "NearestNeighborSearch::get_nearest_neighbors (by index)");
        table_k_ = nb_neighbors;
    }
    const index_t* r = &table_nbr_[size_t(query_point) * table_k_];
    const double* d = &table_d2_[size_t(query_point) * table_k_];
    std::copy(r, r + nb_neighbors, neighbors);
    std::copy(d, d + nb_neighbors, neighbors_sq_dist);
}

index_t NearestNeighborSearch::get_nearest_neighbor(const double* query_point) const {
    index_t result;
    double sq_dist;
    get_nearest_neighbors(1, query_point, &result, &sq_dist);
    return result;
}

void NearestNeighborSearch::get_nearest_neighbors_all(index_t nb_neighbors, index_t* neighbors,
                                                      double* sq_dist) const {
    std::lock_guard<std::mutex> g(lock_);
    check(b200vor_knn(ctx_, nb_neighbors, neighbors, sq_dist, 0), "NearestNeighborSearch::get_nearest_neighbors_all");
}

void NearestNeighborSearch::get_nearest_neighbors_batch(index_t nb_queries, const double* query_points,
                                                        index_t nb_neighbors, index_t* neighbors,
                                                        double* sq_dist) const {
    std::lock_guard<std::mutex> g(lock_);
    check(b200vor_knn_points(ctx_, nb_queries, query_points, nb_neighbors, neighbors, sq_dist, 0),
          "NearestNeighborSearch::get_nearest_neighbors_batch");
}

// ------------------------------------------------------------------ Delaunay_NearestNeighbors
Delaunay_NearestNeighbors::Delaunay_NearestNeighbors(coord_index_t dimension)
    : NN_(nullptr), nb_vertices_(0), vertices_(nullptr), default_nb_neighbors_(20) {
    if(dimension != 3) {
        // Delaunay::InvalidDimension (delaunay.cpp:92-106)
        throw std::runtime_error("Invalid dimension: dimension = " + std::to_string(int(dimension)) +
                                 " is not supported by the B200 Delaunay_NearestNeighbors implementation");
    }
    NN_ = NearestNeighborSearch::create(dimension, "B200NN");
}

Delaunay_NearestNeighbors::~Delaunay_NearestNeighbors() { delete NN_; }

void Delaunay_NearestNeighbors::store_neighbors(index_t v, index_t nb, const index_t* row, const double* d2) {
    // delaunay_nn.cpp:104-145: row holds nb+1 entries, nearest first (normally v itself)
    std::vector<index_t>& out = neighbors_[v];
    out.clear();
    for(index_t k = 0; k < nb + 1; ++k) {
        index_t j = row[k];
        if(j == index_t(-1)) {
            break;
        }
        if(d2[k] == 0.0) {
            if(j < v) {
                out.clear();   // v duplicates a lower index: it owns no cell and no neighbours
                return;
            }
            continue;          // v itself, or a duplicate with a higher index
        }
        out.push_back(j);
    }
    if(out.size() > nb) {
        out.resize(nb);
    }
}

void Delaunay_NearestNeighbors::set_vertices(index_t nb_vertices, const double* vertices) {
    nb_vertices_ = nb_vertices;
    vertices_ = vertices;
    NN_->set_points(nb_vertices, vertices);
    index_t nb = std::min(default_nb_neighbors_, nb_vertices - 1);
    std::vector<index_t> nbr(size_t(nb_vertices) * (nb + 1));
    std::vector<double> d2(size_t(nb_vertices) * (nb + 1));
    NN_->get_nearest_neighbors_all(nb + 1, nbr.data(), d2.data());
    neighbors_.assign(nb_vertices, std::vector<index_t>());
    for(index_t v = 0; v < nb_vertices; ++v) {
        store_neighbors(v, nb, &nbr[size_t(v) * (nb + 1)], &d2[size_t(v) * (nb + 1)]);
    }
}

index_t Delaunay_NearestNeighbors::nearest_vertex(const double* p) const {
    return NN_->get_nearest_neighbor(p);
}

void Delaunay_NearestNeighbors::get_neighbors(index_t v, std::vector<index_t>& neighbors) const {
    neighbors = neighbors_[v];
}

void Delaunay_NearestNeighbors::enlarge_neighborhood(index_t v, index_t nb) {
    nb = std::min(nb, nb_vertices_ - 1);
    std::vector<index_t> row(nb + 1);
    std::vector<double> d2(nb + 1);
    NN_->get_nearest_neighbors(nb + 1, vertex_ptr(v), row.data(), d2.data());
    store_neighbors(v, nb, row.data(), d2.data());
}

// ------------------------------------------------------------------ ConvexCellBatch
ConvexCellBatch::ConvexCellBatch() : ctx_(new_context()) {}

ConvexCellBatch::~ConvexCellBatch() { b200vor_destroy(ctx_); }

void ConvexCellBatch::init_with_box(double xmin, double ymin, double zmin, double xmax, double ymax,
                                    double zmax) {
    const double box[6] = {xmin, ymin, zmin, xmax, ymax, zmax};
    check(b200vor_set_domain_box(ctx_, box), "ConvexCellBatch::init_with_box");
}

void ConvexCellBatch::set_periodic(double period) {
    check(b200vor_set_domain_periodic(ctx_, period), "ConvexCellBatch::set_periodic");
}

void ConvexCellBatch::compute(index_t nb_seeds, const double* xyz, const double* weights, bool with_faces,
                              bool with_vertices) {
    check(b200vor_set_points(ctx_, nb_seeds, xyz, 3, weights, 0), "ConvexCellBatch::compute (set_points)");
    b200vor_cell_params p;
    memset(&p, 0, sizeof(p));
    with_faces = with_faces || with_vertices;
    p.want_faces = with_vertices ? 2 : (with_faces ? 1 : 0);
    check(b200vor_compute_cells(ctx_, &p), "ConvexCellBatch::compute");
    m_.resize(nb_seeds);
    mg_.resize(3 * size_t(nb_seeds));
    status_.resize(nb_seeds);
    check(b200vor_get_cells(ctx_, m_.data(), mg_.data(), status_.data(), 0), "ConvexCellBatch::compute (get_cells)");
    if(with_faces) {
        nfaces_.resize(nb_seeds);
        face_nbr_.resize(size_t(nb_seeds) * 64);
        face_area_.resize(size_t(nb_seeds) * 64);
        check(b200vor_get_faces(ctx_, nfaces_.data(), face_nbr_.data(), face_area_.data(), 64, 0),
              "ConvexCellBatch::compute (get_faces)");
    } else {
        nfaces_.clear();
        face_nbr_.clear();
        face_area_.clear();
    }
    if(with_vertices) {
        nverts_.resize(nb_seeds);
        vert_xyz_.resize(size_t(nb_seeds) * 64 * 3);
        vert_faces_.resize(size_t(nb_seeds) * 64);
        check(b200vor_get_vertices(ctx_, nverts_.data(), vert_xyz_.data(), vert_faces_.data(), 64, 0),
              "ConvexCellBatch::compute (get_vertices)");
    } else {
        nverts_.clear();
        vert_xyz_.clear();
        vert_faces_.clear();
    }
}

index_t ConvexCellBatch::nb_t(index_t c) const {
    if(nverts_.empty() || nverts_[c] == 0xFFFFFFFFu || nverts_[c] > 64) {
        return 0;   // not available (kNN-table pipeline) or more vertices than the list holds
    }
    return nverts_[c];
}

vec3 ConvexCellBatch::triangle_point(index_t c, index_t t) const {
    const double* p = &vert_xyz_[(size_t(c) * 64 + t) * 3];
    vec3 r = {p[0], p[1], p[2]};
    return r;
}

index_t ConvexCellBatch::triangle_v_local_index(index_t c, index_t t, index_t lv) const {
    return (vert_faces_[size_t(c) * 64 + t] >> (8 * lv)) & 0xFFu;
}

void ConvexCellBatch::facet_loops(index_t c, std::vector<std::vector<index_t> >& loops) const {
    const index_t nt = nb_t(c), nf = nb_v(c);
    loops.assign(nf, std::vector<index_t>());
    // per face: for every vertex on it, (vertex, face it is entered through, face it is left through);
    // all triples have the same orientation, so the vertex after (v, g, h) is the one entered through h
    std::vector<std::vector<index_t> > ring(nf);
    for(index_t t = 0; t < nt; ++t) {
        for(index_t l = 0; l < 3; ++l) {
            ring[triangle_v_local_index(c, t, l)].push_back(t * 3 + l);
        }
    }
    for(index_t f = 0; f < nf; ++f) {
        const std::vector<index_t>& R = ring[f];
        if(R.empty()) {
            continue;
        }
        std::vector<index_t>& L = loops[f];
        index_t cur = R[0];
        for(size_t step = 0; step < R.size(); ++step) {
            const index_t t = cur / 3, l = cur % 3;
            L.push_back(t);
            const index_t h = triangle_v_local_index(c, t, (l + 2) % 3);
            bool found = false;
            for(size_t k = 0; k < R.size() && !found; ++k) {
                const index_t t2 = R[k] / 3, l2 = R[k] % 3;
                if(t2 != t && triangle_v_local_index(c, t2, (l2 + 1) % 3) == h) {
                    cur = R[k];
                    found = true;
                }
            }
            if(!found || cur == R[0]) {
                break;
            }
        }
    }
}

index_t ConvexCellBatch::save(index_t c, std::ostream& out, global_index_t v_offset, double shrink) const {
    const index_t nt = nb_t(c);
    if(nt == 0) {
        return 0;
    }
    vec3 g = {0.0, 0.0, 0.0};
    if(shrink != 0.0) {
        g = barycenter(c);
    }
    out.precision(17);
    for(index_t t = 0; t < nt; ++t) {
        vec3 p = triangle_point(c, t);
        if(shrink != 0.0) {
            p.x = shrink * g.x + (1.0 - shrink) * p.x;
            p.y = shrink * g.y + (1.0 - shrink) * p.y;
            p.z = shrink * g.z + (1.0 - shrink) * p.z;
        }
        out << "v " << p.x << " " << p.y << " " << p.z << std::endl;
    }
    std::vector<std::vector<index_t> > loops;
    facet_loops(c, loops);
    for(size_t f = 0; f < loops.size(); ++f) {
        if(loops[f].size() < 3) {
            continue;
        }
        out << "f ";
        for(size_t k = 0; k < loops[f].size(); ++k) {
            out << (loops[f][k] + v_offset) << " ";
        }
        out << std::endl;
    }
    return nt;
}

void ConvexCellBatch::save_cells(const std::string& filename, double shrink) const {
    std::ofstream out(filename.c_str());
    if(!out) {
        throw std::runtime_error("ConvexCellBatch::save_cells: cannot open " + filename);
    }
    global_index_t v_off = 1;
    for(index_t c = 0; c < nb_cells(); ++c) {
        v_off += save(c, out, v_off, shrink);
    }
}

index_t ConvexCellBatch::delaunay_tetrahedra(std::vector<global_index_t>& tets) const {
    tets.clear();
    const index_t n = nb_cells();
    index_t unavailable = 0;
    for(index_t c = 0; c < n; ++c) {
        // no vertex list (cell finished by the kNN-table pipeline, or more than 64 vertices): its
        // tetrahedra are missing from the result -- counted, never silently dropped
        if(!nverts_.empty() && (nverts_[c] == 0xFFFFFFFFu || nverts_[c] > 64u)) {
            ++unavailable;
            continue;
        }
        const index_t nt = nb_t(c);
        for(index_t t = 0; t < nt; ++t) {
            global_index_t v[3];
            bool keep = true;
            for(index_t l = 0; l < 3 && keep; ++l) {
                v[l] = v_global_index(c, triangle_v_local_index(c, t, l));
                keep = v[l] < DOMAIN_ID && c < v[l] % n;   // reported by the smallest real index
            }
            if(keep) {
                tets.push_back(c);
                tets.push_back(v[0]);
                tets.push_back(v[1]);
                tets.push_back(v[2]);
            }
        }
    }
    return unavailable;
}

void ConvexCellBatch::compute_mg(index_t c, double& m, vec3& mg) const {
    m = m_[c];
    mg.x = mg_[3 * size_t(c)];
    mg.y = mg_[3 * size_t(c) + 1];
    mg.z = mg_[3 * size_t(c) + 2];
}

vec3 ConvexCellBatch::barycenter(index_t c) const {
    // ConvexCell::barycenter: mg / m (convex_cell.cpp:1255-1263)
    vec3 r;
    double s = 1.0 / m_[c];
    r.x = s * mg_[3 * size_t(c)];
    r.y = s * mg_[3 * size_t(c) + 1];
    r.z = s * mg_[3 * size_t(c) + 2];
    return r;
}

double ConvexCellBatch::total_volume() const {
    double s = 0.0;
    for(size_t i = 0; i < m_.size(); ++i) {
        s += m_[i];
    }
    return s;
}

namespace {
// integer translation of a periodic instance and back (delaunay/periodic.h:104-113, periodic.cpp:80-84)
void instance_translation(uint32_t inst, int t[3]) {
    const int i = inst == 0 ? 13 : (inst <= 13 ? int(inst) - 1 : int(inst));
    t[0] = i / 9 - 1;
    t[1] = (i / 3) % 3 - 1;
    t[2] = i % 3 - 1;
}
uint32_t translation_instance(const int t[3]) {
    const int i = (t[2] + 1) + 3 * (t[1] + 1) + 9 * (t[0] + 1);
    return uint32_t(i == 13 ? 0 : (i < 13 ? i + 1 : i));
}
}  // namespace

void ConvexCellBatch::tet_adjacency(const std::vector<global_index_t>& tets, std::vector<int64_t>& adjacent,
                                    index_t nb_real_vertices) {
    const size_t nt = tets.size() / 4;
    adjacent.assign(4 * nt, -1);
    struct Facet {
        global_index_t v[3];
        uint32_t owner;   // 4 c + lf
    };
    std::vector<Facet> F(4 * nt);
    for(size_t c = 0; c < nt; ++c) {
        for(uint32_t lf = 0; lf < 4; ++lf) {
            Facet& f = F[4 * c + lf];
            uint32_t k = 0;
            for(uint32_t lv = 0; lv < 4; ++lv) {
                if(lv != lf) {
                    f.v[k++] = tets[4 * c + lv];
                }
            }
            if(nb_real_vertices != 0) {
                // periodic: the same facet carries different instance numbers in the frames of its two
                // tetrahedra; bring it to the frame where its smallest (real, translation) vertex is
                // instance 0
                const index_t n = nb_real_vertices;
                int tr[3][3];
                int anchor = 0;
                for(int a = 0; a < 3; ++a) {
                    instance_translation(f.v[a] / n, tr[a]);
                    const global_index_t ra = f.v[a] % n, r0 = f.v[anchor] % n;
                    if(ra < r0 || (ra == r0 && std::lexicographical_compare(tr[a], tr[a] + 3, tr[anchor], tr[anchor] + 3))) {
                        anchor = a;
                    }
                }
                bool representable = true;
                global_index_t canon[3];
                for(int a = 0; a < 3; ++a) {
                    int rel[3];
                    for(int d = 0; d < 3; ++d) {
                        rel[d] = tr[a][d] - tr[anchor][d];
                        representable = representable && rel[d] >= -1 && rel[d] <= 1;
                    }
                    canon[a] = representable ? f.v[a] % n + n * translation_instance(rel) : 0;
                }
                if(representable) {
                    f.v[0] = canon[0]; f.v[1] = canon[1]; f.v[2] = canon[2];
                }
            }
            std::sort(f.v, f.v + 3);
            f.owner = uint32_t(4 * c + lf);
        }
    }
    std::sort(F.begin(), F.end(), [](const Facet& a, const Facet& b) {
        if(a.v[0] != b.v[0]) return a.v[0] < b.v[0];
        if(a.v[1] != b.v[1]) return a.v[1] < b.v[1];
        if(a.v[2] != b.v[2]) return a.v[2] < b.v[2];
        return a.owner < b.owner;
    });
    for(size_t k = 0; k + 1 < F.size(); ++k) {
        const Facet &a = F[k], &b = F[k + 1];
        if(a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2]) {
            adjacent[a.owner] = int64_t(b.owner / 4);
            adjacent[b.owner] = int64_t(a.owner / 4);
            ++k;
        }
    }
}

// ------------------------------------------------------------------ CentroidalVoronoiTesselation
CentroidalVoronoiTesselation::CentroidalVoronoiTesselation(coord_index_t dimension) : ctx_(nullptr) {
    if(dimension != 3) {
        throw std::runtime_error("CentroidalVoronoiTesselation: only dimension 3 is implemented");
    }
    ctx_ = new_context();
}

CentroidalVoronoiTesselation::~CentroidalVoronoiTesselation() { b200vor_destroy(ctx_); }

void CentroidalVoronoiTesselation::set_domain_box(const double* box6) {
    check(b200vor_set_domain_box(ctx_, box6), "CentroidalVoronoiTesselation::set_domain_box");
}

void CentroidalVoronoiTesselation::set_periodic(double period) {
    check(b200vor_set_domain_periodic(ctx_, period), "CentroidalVoronoiTesselation::set_periodic");
}

void CentroidalVoronoiTesselation::set_volumetric_mesh(index_t nb_vertices, const double* vertices, index_t nb_tets,
                                                       const index_t* tets) {
    check(b200vor_set_domain_tets(ctx_, nb_vertices, vertices, nb_tets, tets),
          "CentroidalVoronoiTesselation::set_volumetric_mesh");
}

void CentroidalVoronoiTesselation::set_domain_planes(index_t nb_planes, const double* planes) {
    check(b200vor_set_domain_planes(ctx_, nb_planes, planes), "CentroidalVoronoiTesselation::set_domain_planes");
}

void CentroidalVoronoiTesselation::set_points(index_t nb_points, const double* points) {
    points_.assign(points, points + 3 * size_t(nb_points));
    point_is_locked_.clear();
}

void CentroidalVoronoiTesselation::lock_point(index_t i) {
    if(point_is_locked_.empty()) {
        point_is_locked_.assign(nb_points(), false);
    }
    point_is_locked_[i] = true;
}

void CentroidalVoronoiTesselation::funcgrad(index_t n, double* x, double& f, double* g) {
    const index_t nb = n / 3;
    check(b200vor_set_points(ctx_, nb, x, 3, nullptr, 0), "funcgrad (set_points)");
    b200vor_cell_params p;
    memset(&p, 0, sizeof(p));
    p.want_energy = 1;
    check(b200vor_compute_cells(ctx_, &p), "funcgrad (compute_cells)");
    check(b200vor_get_cvt(ctx_, nullptr, g, &f, 0), "funcgrad (get_cvt)");
    if(!point_is_locked_.empty()) {
        for(index_t i = 0; i < nb && i < index_t(point_is_locked_.size()); ++i) {
            if(point_is_locked_[i]) {
                g[3 * size_t(i)] = g[3 * size_t(i) + 1] = g[3 * size_t(i) + 2] = 0.0;
            }
        }
    }
}

void CentroidalVoronoiTesselation::Newton_iterations(index_t nb_iter, index_t m) {
    const size_t N = points_.size();
    if(N == 0 || nb_iter == 0) {
        return;
    }
    m = std::max<index_t>(m, 1);
    std::vector<double> x(points_), g(N), xn(N), gn(N), d(N);
    std::vector<std::vector<double> > S, Y;
    std::vector<double> rho;
    double f = 0.0;
    funcgrad(index_t(N), x.data(), f, g.data());
    auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) {
        double s = 0.0;
        for(size_t k = 0; k < N; ++k) {
            s += a[k] * b[k];
        }
        return s;
    };
    for(index_t it = 0; it < nb_iter; ++it) {
        // two-loop recursion: d = -H g
        d = g;
        std::vector<double> al(S.size());
        for(size_t k = S.size(); k-- > 0;) {
            al[k] = rho[k] * dot(S[k], d);
            for(size_t j = 0; j < N; ++j) {
                d[j] -= al[k] * Y[k][j];
            }
        }
        if(!S.empty()) {
            const double gamma = dot(S.back(), Y.back()) / dot(Y.back(), Y.back());
            for(size_t j = 0; j < N; ++j) {
                d[j] *= gamma;
            }
        } else {
            // first step: the Lloyd step x - g / (2 m_i) is exact for fixed cells; 1/(2 <m>) per point
            const double scale = double(nb_points()) / 2.0;
            for(size_t j = 0; j < N; ++j) {
                d[j] *= scale;
            }
        }
        for(size_t k = 0; k < S.size(); ++k) {
            const double be = rho[k] * dot(Y[k], d);
            for(size_t j = 0; j < N; ++j) {
                d[j] += (al[k] - be) * S[k][j];
            }
        }
        for(size_t j = 0; j < N; ++j) {
            d[j] = -d[j];
        }
        double gd = dot(g, d);
        if(!(gd < 0.0)) {   // not a descent direction: restart from steepest descent
            S.clear(); Y.clear(); rho.clear();
            for(size_t j = 0; j < N; ++j) {
                d[j] = -g[j] * double(nb_points()) / 2.0;
            }
            gd = dot(g, d);
        }
        // Armijo backtracking
        double t = 1.0, fn = f;
        bool ok = false;
        for(int ls = 0; ls < 20; ++ls) {
            for(size_t j = 0; j < N; ++j) {
                xn[j] = x[j] + t * d[j];
            }
            funcgrad(index_t(N), xn.data(), fn, gn.data());
            if(fn <= f + 1e-4 * t * gd) {
                ok = true;
                break;
            }
            t *= 0.5;
        }
        if(!ok) {
            break;
        }
        std::vector<double> s(N), y(N);
        for(size_t j = 0; j < N; ++j) {
            s[j] = xn[j] - x[j];
            y[j] = gn[j] - g[j];
        }
        const double sy = dot(s, y);
        if(sy > 1e-300) {
            if(S.size() == m) {
                S.erase(S.begin()); Y.erase(Y.begin()); rho.erase(rho.begin());
            }
            S.push_back(s); Y.push_back(y); rho.push_back(1.0 / sy);
        }
        x.swap(xn);
        g.swap(gn);
        f = fn;
    }
    points_ = x;
    last_f_ = f;
}

void CentroidalVoronoiTesselation::Lloyd_iterations(index_t nb_iter) {
    const index_t n = nb_points();
    if(n == 0 || nb_iter == 0) {
        return;
    }
    check(b200vor_set_points(ctx_, n, points_.data(), 3, nullptr, 0), "Lloyd_iterations (set_points)");
    if(point_is_locked_.empty()) {
        // all iterations stay on the device
        check(b200vor_lloyd(ctx_, nb_iter, nullptr, points_.data(), 0), "Lloyd_iterations");
        return;
    }
    std::vector<double> next(points_.size());
    for(index_t it = 0; it < nb_iter; ++it) {
        check(b200vor_lloyd(ctx_, 1, nullptr, next.data(), 0), "Lloyd_iterations");
        for(index_t j = 0; j < n; ++j) {
            if(!point_is_locked_[j]) {
                for(int c = 0; c < 3; ++c) {
                    points_[3 * size_t(j) + c] = next[3 * size_t(j) + c];
                }
            }
        }
        check(b200vor_set_points(ctx_, n, points_.data(), 3, nullptr, 0), "Lloyd_iterations (set_points)");
    }
}

// ------------------------------------------------------------------ OptimalTransportMap3d
OptimalTransportMap3d::OptimalTransportMap3d()
    : ctx_(new_context()), epsilon_(0.01), converged_(false), newton_iters_(0), period_(0.0), has_box_(false) {
    for(int i = 0; i < 6; ++i) {
        box_[i] = (i < 3) ? 0.0 : 1.0;
    }
}

OptimalTransportMap3d::~OptimalTransportMap3d() { b200vor_destroy(ctx_); }

void OptimalTransportMap3d::set_domain_box(const double* box6) {
    check(b200vor_set_domain_box(ctx_, box6), "OptimalTransportMap3d::set_domain_box");
    std::copy(box6, box6 + 6, box_);
    has_box_ = true;
    period_ = 0.0;
}

void OptimalTransportMap3d::set_periodic(double period) {
    check(b200vor_set_domain_periodic(ctx_, period), "OptimalTransportMap3d::set_periodic");
    period_ = period;
}

void OptimalTransportMap3d::set_points(index_t nb_points, const double* points) {
    points_.assign(points, points + 3 * size_t(nb_points));
    weights_.assign(nb_points, 0.0);
    nu_.clear();
    level_iters_.clear();
    converged_ = false;
}

void OptimalTransportMap3d::set_nu(index_t i, double nu) {
    if(nu_.empty()) {
        nu_.assign(weights_.size(), -1.0);
    }
    nu_[i] = nu;
}

// Newton solve on the points [0, e) at the current weights_[0, e); target masses: nu_ rescaled so
// that the prefix carries the whole measure (uniform when nu_ is not set)
void OptimalTransportMap3d::solve_prefix(index_t e, index_t max_iterations, const char* who) {
    check(b200vor_set_points(ctx_, e, points_.data(), 3, nullptr, 0), who);
    std::vector<double> nu_level;
    if(!nu_.empty()) {
        double all = 0.0, part = 0.0;
        for(size_t i = 0; i < nu_.size(); ++i) {
            all += nu_[i];
            if(i < e) {
                part += nu_[i];
            }
        }
        nu_level.assign(nu_.begin(), nu_.begin() + e);
        const double s = all / part;
        if(e != nu_.size()) {
            for(index_t i = 0; i < e; ++i) {
                nu_level[i] *= s;
            }
        }
    }
    b200vor_otm_params p;
    memset(&p, 0, sizeof(p));
    p.max_newton = max_iterations;
    p.epsilon = epsilon_;
    b200vor_otm_result r;
    // The damped Newton method starts from weights whose cells are all non-empty and certified.
    // Extrapolated start weights (optimize_levels) can miss that for a few cells: the weights are then
    // pulled towards the Voronoi diagram (all weights equal: every cell of distinct points is
    // non-empty) by halving.
    for(int attempt = 0;; ++attempt) {
        const int rc = b200vor_otm_newton(ctx_, nu_level.empty() ? nullptr : nu_level.data(), weights_.data(), &p, &r, 0);
        if(rc == B200VOR_OK) {
            break;
        }
        const char* msg = b200vor_last_error();
        if(attempt < 8 && msg && (strstr(msg, "an initial cell is empty") ||
                                  strstr(msg, "of the start weights are flagged"))) {
            for(index_t i = 0; i < e; ++i) {
                weights_[i] *= 0.5;
            }
            continue;
        }
        check(rc, who);
    }
    converged_ = r.converged != 0;
    newton_iters_ = r.newton_iters;
}

void OptimalTransportMap3d::optimize(index_t max_iterations) {
    const index_t n = nb_points();
    if(n == 0) {
        return;
    }
    for(size_t i = 0; i < nu_.size(); ++i) {
        if(!(nu_[i] > 0.0)) {
            throw std::runtime_error("OptimalTransportMap3d::optimize: set_nu must be called for every point or for none");
        }
    }
    solve_prefix(n, max_iterations, "OptimalTransportMap3d::optimize");
    measure_.resize(n);
    mg_.resize(3 * size_t(n));
    check(b200vor_get_cells(ctx_, measure_.data(), mg_.data(), nullptr, 0), "OptimalTransportMap3d::optimize (get_cells)");
}

// value at q of the least-squares plane through (p_j, w_j), j in row[0..k): normal equations in the
// frame of q (4 unknowns: value at q and gradient), Gaussian elimination with partial pivoting;
// a rank-deficient neighbourhood (k < 4, coplanar points) falls back to the nearest weight
double OptimalTransportMap3d::extrapolate_weight(const double* pts, const double* w, const double* q,
                                                 const uint32_t* row, uint32_t k) {
    if(k < 4) {
        return w[row[0]];
    }
    double A[4][5] = {};
    for(uint32_t t = 0; t < k; ++t) {
        const double* p = pts + 3 * size_t(row[t]);
        const double f[4] = {1.0, p[0] - q[0], p[1] - q[1], p[2] - q[2]};
        for(int r = 0; r < 4; ++r) {
            for(int c = 0; c < 4; ++c) {
                A[r][c] += f[r] * f[c];
            }
            A[r][4] += f[r] * w[row[t]];
        }
    }
    const double scale = A[1][1] + A[2][2] + A[3][3];
    for(int c = 0; c < 4; ++c) {
        int piv = c;
        for(int r = c + 1; r < 4; ++r) {
            if(fabs(A[r][c]) > fabs(A[piv][c])) {
                piv = r;
            }
        }
        if(!(fabs(A[piv][c]) > 1e-10 * (c == 0 ? double(k) : scale))) {
            return w[row[0]];
        }
        for(int j = 0; j < 5; ++j) {
            std::swap(A[c][j], A[piv][j]);
        }
        for(int r = c + 1; r < 4; ++r) {
            const double m = A[r][c] / A[c][c];
            for(int j = c; j < 5; ++j) {
                A[r][j] -= m * A[c][j];
            }
        }
    }
    // back substitution; only the first unknown (the value at q) is wanted, but it depends on all
    double x[4];
    for(int r = 3; r >= 0; --r) {
        double v = A[r][4];
        for(int j = r + 1; j < 4; ++j) {
            v -= A[r][j] * x[j];
        }
        x[r] = v / A[r][r];
    }
    // no overshoot: outside the hull of its neighbours (new points next to the domain boundary) the
    // fit EXTRApolates; the start weight stays within the range of the weights it was fitted to
    double lo = w[row[0]], hi = w[row[0]];
    for(uint32_t t = 1; t < k; ++t) {
        lo = std::min(lo, w[row[t]]);
        hi = std::max(hi, w[row[t]]);
    }
    return std::min(std::max(x[0], lo), hi);
}

void OptimalTransportMap3d::optimize_levels(const std::vector<index_t>& levels, index_t max_iterations) {
    const char* who = "OptimalTransportMap3d::optimize_levels";
    const index_t n = nb_points();
    if(n == 0) {
        return;
    }
    if(levels.size() < 2 || levels.front() != 0 || levels.back() != n) {
        throw std::runtime_error(std::string(who) + ": levels must run from 0 to nb_points");
    }
    for(size_t l = 0; l + 1 < levels.size(); ++l) {
        if(levels[l + 1] <= levels[l]) {
            throw std::runtime_error(std::string(who) + ": levels must increase");
        }
    }
    for(size_t i = 0; i < nu_.size(); ++i) {
        if(!(nu_[i] > 0.0)) {
            throw std::runtime_error(std::string(who) + ": set_nu must be called for every point or for none");
        }
    }
    level_iters_.clear();
    std::vector<uint32_t> nearest;
    std::vector<double> solved;
    for(size_t l = 0; l + 1 < levels.size(); ++l) {
        const index_t b = levels[l], e = levels[l + 1];
        if(b != 0) {
            // the context still holds the points [0, b) of the level below: one batched by-point kNN
            // query on the device.  It is a box query: on a periodic cube it runs on the cube without
            // the wrap (the rows feed a start value, not a result).
            if(period_ > 0.0) {
                const double cube[6] = {0.0, 0.0, 0.0, period_, period_, period_};
                check(b200vor_set_domain_box(ctx_, cube), who);
            }
            // a LINEAR fit of the solved weights over the k nearest solved points, evaluated at EVERY point
            // of the new level, the solved ones included: what carries over between levels is the smooth
            // part of the potential.  (A nearest-weight copy jumps by |grad w| x (coarse spacing) between
            // neighbouring new points; the per-point part of the solved weights compensates the cell
            // irregularities of one level's point set, order (coarse spacing)^2: both empty cells.)
            const uint32_t k = std::min<index_t>(b, 24);
            nearest.resize(size_t(e) * k);
            check(b200vor_knn_points(ctx_, e, points_.data(), k, nearest.data(), nullptr, 0), who);
            if(period_ > 0.0) {
                check(b200vor_set_domain_periodic(ctx_, period_), who);
            }
            solved.assign(weights_.begin(), weights_.begin() + b);
            for(index_t i = 0; i < e; ++i) {
                const uint32_t* row = nearest.data() + size_t(i) * k;
                weights_[i] = extrapolate_weight(points_.data(), solved.data(), points_.data() + 3 * size_t(i), row, k);
            }
        }
        try {
            solve_prefix(e, max_iterations, who);
        } catch(const std::runtime_error& err) {
            throw std::runtime_error(std::string(err.what()) + " [level " + std::to_string(l) + ": points 0.." +
                                     std::to_string(e) + "]");
        }
        level_iters_.push_back(newton_iters_);
    }
    measure_.resize(n);
    mg_.resize(3 * size_t(n));
    check(b200vor_get_cells(ctx_, measure_.data(), mg_.data(), nullptr, 0), who);
}

vec3 OptimalTransportMap3d::cell_centroid(index_t i) const {
    vec3 r;
    double s = 1.0 / measure_[i];
    r.x = s * mg_[3 * size_t(i)];
    r.y = s * mg_[3 * size_t(i) + 1];
    r.z = s * mg_[3 * size_t(i) + 2];
    return r;
}

// ------------------------------------------------------------------ BRIO order
uint64_t hilbert_index_3d(uint32_t x, uint32_t y, uint32_t z, int bits) {
    uint32_t X[3] = {x, y, z};
    // axes -> transpose (J. Skilling, "Programming the Hilbert curve", 2004)
    for(uint32_t Q = 1u << (bits - 1); Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for(int i = 0; i < 3; ++i) {
            if(X[i] & Q) {
                X[0] ^= P;
            } else {
                const uint32_t t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for(uint32_t Q = 1u << (bits - 1); Q > 1; Q >>= 1) {
        if(X[2] & Q) {
            t ^= Q - 1;
        }
    }
    uint64_t key = 0;
    for(int i = 0; i < 3; ++i) {
        X[i] ^= t;
    }
    for(int b = 0; b < bits; ++b) {
        for(int i = 0; i < 3; ++i) {
            key |= uint64_t((X[i] >> b) & 1u) << (3 * b + (2 - i));
        }
    }
    return key;
}

void compute_BRIO_order(index_t nb_vertices, const double* vertices, std::vector<index_t>& sorted_indices,
                        index_t stride, index_t threshold, double ratio, std::vector<index_t>* levels,
                        uint64_t seed) {
    const index_t n = nb_vertices;
    sorted_indices.resize(n);
    for(index_t i = 0; i < n; ++i) {
        sorted_indices[i] = i;
    }
    if(levels) {
        levels->clear();
        levels->push_back(0);
    }
    if(n == 0) {
        return;
    }
    // Fisher-Yates with a 64-bit LCG (the reference uses GEO::random_shuffle: another stream of numbers)
    uint64_t state = seed * 6364136223846793005ull + 1442695040888963407ull;
    for(index_t i = n - 1; i > 0; --i) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const index_t j = index_t((state >> 33) % uint64_t(i + 1));
        std::swap(sorted_indices[i], sorted_indices[j]);
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for(index_t i = 0; i < n; ++i) {
        for(int d = 0; d < 3; ++d) {
            lo[d] = std::min(lo[d], vertices[size_t(i) * stride + d]);
            hi[d] = std::max(hi[d], vertices[size_t(i) * stride + d]);
        }
    }
    const int bits = 16;
    double scale[3];
    for(int d = 0; d < 3; ++d) {
        scale[d] = hi[d] > lo[d] ? double(1u << bits) / (hi[d] - lo[d]) : 0.0;
    }
    // the nested intervals from the outside in (mesh_reorder.cpp:1126-1136): [m, e) is sorted, [b, m) recursed
    std::vector<std::pair<index_t, index_t> > bounds;
    index_t e = n;
    for(;;) {
        index_t m = 0;
        if(e > threshold) {
            m = index_t(double(e) * ratio);
        }
        bounds.push_back(std::make_pair(m, e));
        if(m == 0) {
            break;
        }
        e = m;
    }
    std::vector<std::pair<uint64_t, index_t> > keyed;
    for(size_t l = bounds.size(); l-- > 0;) {
        const index_t m = bounds[l].first, end = bounds[l].second;
        keyed.resize(end - m);
        for(index_t i = m; i < end; ++i) {
            const double* p = vertices + size_t(sorted_indices[i]) * stride;
            uint32_t q[3];
            for(int d = 0; d < 3; ++d) {
                const double v = (p[d] - lo[d]) * scale[d];
                q[d] = uint32_t(std::min(std::max(v, 0.0), double((1u << bits) - 1)));
            }
            keyed[i - m] = std::make_pair(hilbert_index_3d(q[0], q[1], q[2], bits), sorted_indices[i]);
        }
        std::sort(keyed.begin(), keyed.end());
        for(index_t i = m; i < end; ++i) {
            sorted_indices[i] = keyed[i - m].second;
        }
        if(levels) {
            levels->push_back(end);
        }
    }
}

}  // namespace B200GEO
