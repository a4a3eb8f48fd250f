"""The C++ host side (geogram_b200/host: mirrors of GEO::NearestNeighborSearch,
GEO::Delaunay_NearestNeighbors, VBW::ConvexCell queries, CVT Lloyd) above the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "geogram_b200", "lib")


def _build():
    """(re)builds the libraries when the CUDA toolchain is here; a box without nvcc uses the
    prebuilt ones that travel with the snapshot, and skips if there are none"""
    import shutil
    have_nvcc = shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")
    if have_nvcc:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "geogram_b200", "csrc")])
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "geogram_b200", "host")])
    elif not (os.path.exists(os.path.join(LIB, "libb200vor.so")) and os.path.exists(os.path.join(LIB, "libb200geo.so"))):
        pytest.skip("no CUDA toolkit and no prebuilt geogram_b200/lib/*.so")


def test_host_library_exports_the_mirrored_interfaces():
    _build()
    out = subprocess.check_output(["nm", "-D", "--demangle", os.path.join(LIB, "libb200geo.so")]).decode()
    for sym in [
        "B200GEO::NearestNeighborSearch::create(",
        "B200GEO::NearestNeighborSearch::set_points(unsigned int, double const*)",
        "B200GEO::NearestNeighborSearch::set_points(unsigned int, double const*, unsigned int)",
        "B200GEO::NearestNeighborSearch::get_nearest_neighbors(unsigned int, double const*, unsigned int*, double*) const",
        "B200GEO::NearestNeighborSearch::get_nearest_neighbors(unsigned int, unsigned int, unsigned int*, double*) const",
        "B200GEO::Delaunay_NearestNeighbors::set_vertices(unsigned int, double const*)",
        "B200GEO::Delaunay_NearestNeighbors::enlarge_neighborhood(unsigned int, unsigned int)",
        "B200GEO::ConvexCellBatch::compute(",
        "B200GEO::ConvexCellBatch::compute_mg(",
        "B200GEO::CentroidalVoronoiTesselation::Lloyd_iterations(unsigned int)",
        "B200GEO::CentroidalVoronoiTesselation::funcgrad(unsigned int, double*, double&, double*)",
        "B200GEO::CentroidalVoronoiTesselation::Newton_iterations(unsigned int, unsigned int)",
        "B200GEO::ConvexCellBatch::save_cells(",
        "B200GEO::ConvexCellBatch::delaunay_tetrahedra(",
        "B200GEO::OptimalTransportMap3d::set_points(unsigned int, double const*)",
        "B200GEO::OptimalTransportMap3d::optimize(unsigned int)",
        "B200GEO::OptimalTransportMap3d::optimize_levels(",
        "B200GEO::compute_BRIO_order(",
    ]:
        assert sym in out, sym
    # the host library reaches the device only through the C ABI
    undefined = [l.split()[-1] for l in out.splitlines() if " U b200vor_" in l]
    assert "b200vor_create" in undefined and "b200vor_compute_cells" in undefined
    assert "cuda" not in " ".join(undefined).lower()


def test_host_pure_functions_on_cpu():
    """tet adjacency (Delaunay::cell_adjacent) needs no device"""
    _build()
    out = subprocess.run([os.path.join(LIB, "test_host_cpu")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "host cpu test: ok" in out.stdout, out.stdout + out.stderr


def test_host_fails_loudly_without_gpu():
    """No CPU fallback: constructing a search object without a CUDA device throws."""
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([os.path.join(LIB, "test_host")], capture_output=True, text=True)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout) or "b200vor_create" in (p.stderr + p.stdout)


@pytest.mark.gpu
def test_host_cpp_on_gpu():
    p = subprocess.run([os.path.join(LIB, "test_host")], capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0
    assert "host test: ok" in p.stdout


@pytest.mark.gpu
def test_host_cpp_slab_lloyd_two_processes(tmp_path):
    """Multi-GPU Lloyd driven from plain C++ (b200vor_comm_init + b200vor_slab_*, no torch, no MPI):
    two processes, one GPU each, the NCCL id handed over through a file; every rank's rows must equal
    the single-context run.  Needs two GPUs (skipped on a one-GPU box; run with gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(LIB, "test_host_slab")
    idf = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, str(r), "2", idf], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "host slab test: ok" in o, o[-2000:]
