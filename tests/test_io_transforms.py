"""LAS record coordinates in, writer payloads out.

CPU (-m "not gpu"):  the oracle port against the reference's own io/LASFile.cpp, io/PNTSWriter.cpp and
                     io/LASPersistence.* compiled verbatim into oracle/_ref (live, where it is built) and
                     against the committed digests tests/golden/io_golden.json generated from it.
GPU (-m gpu):        swgpu_index_batch_las / swgpu_get_positions / swgpu_get_payload_pnts /
                     swgpu_get_payload_las through the C ABI against the oracle port and the same digests.
Everything is bit-exact: the positions are IEEE doubles computed with the reference's operation order (no
FMA), the payloads float32 / int32.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_io as gio  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "io_golden.json")))["cases"]
CASE_IDS = ["%s-%s" % (c["las"], "shift" if c["shift"] else "world") for c in GOLDEN]


def check_against_golden(case, g):
    res = case["result"]
    assert gio.sha(case["xyz"]) == g["positions"]
    assert case["xyz"][:6].tolist() == g["first_positions"]
    assert (len(res.nodes), len(res.ids)) == (g["nodes"], g["ids"])
    assert gio.sha(gio.canonical_payload(res, case["pnts"])) == g["pnts"]
    assert gio.sha(gio.canonical_payload(res, case["las_out"])) == g["las_records"]
    assert gio.sha(gio.canonical_headers(res, case["headers"])) == g["las_headers"]


# ---------------------------------------------------------------------------------------------------
# oracle pinning (CPU)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("g", GOLDEN, ids=CASE_IDS)
def test_port_equals_committed_reference_golden(port_oracle, g):
    check_against_golden(gio.tiling_case(port_oracle, g["las"], g["shift"]), g)


@pytest.mark.parametrize("shift", [False, True])
def test_port_equals_reference_las_positions(port_oracle, ref_oracle, shift):
    """position_from_las_point on arbitrary int32 records, including the clamped ones."""
    from oracle import sworacle
    rng = np.random.default_rng(3)
    las = rng.integers(-2**31, 2**31 - 1, size=(50_000, 3), dtype=np.int64).astype(np.int32)
    las[::2] = rng.integers(-40_000, 9_000_000, size=(25_000, 3))
    t = sworacle.make_las_transform([0.001, 0.01, 0.00025], [431000.25, 5.1e6, -12.125], [431000.0, 5.1e6 - 50, -30.0],
                                    [439000.5, 5.1e6 + 88000.75, 2100.0],
                                    [435000.25, 5143975.375, 1035.0] if shift else None)
    assert np.array_equal(port_oracle.las_positions(las, t), ref_oracle.las_positions(las, t))


@pytest.mark.parametrize("name", ["small_object", "continental"])
def test_port_equals_reference_payloads(port_oracle, ref_oracle, name):
    """PositionAttribute::extractFromPoints and LASPersistence::persist_points on a tiled cloud."""
    a = gio.tiling_case(port_oracle, name, True, sampling="RANDOM_GRID", tiling="FAST", max_points=300)
    res = a["result"]
    pnts = ref_oracle.payload_pnts(a["clamped"], res.ids)
    las, headers = ref_oracle.payload_las(a["clamped"], res.ids, res.nodes, a["bounds"])
    assert np.array_equal(a["pnts"].view(np.uint32), pnts.view(np.uint32))
    assert np.array_equal(a["las_out"], las)
    filled = res.nodes["count"] > 0  # persist_points returns early for an empty range: no header is written
    for field in ("offset", "scale", "max"):
        assert np.array_equal(a["headers"][field][filled], headers[field][filled]), field
    assert np.array_equal(a["headers"]["scale"], headers["scale"])


def test_las_quantisation_rounds_half_away_from_zero(port_oracle):
    """LASzip's I32_QUANTIZE: (n >= 0) ? (I32)(n + 0.5) : (I32)(n - 0.5) on exact halves."""
    from oracle import sworacle
    nodes = np.zeros(1, sworacle.NODE_DTYPE)
    nodes["count"] = 4
    bounds = (np.array([0.0, 0.0, 0.0]), np.array([64.0, 64.0, 64.0]))  # diagonal > 1 -> scale 0.001
    xyz = np.array([[0.0005, 0.0015, 0.0025], [0.0004999, 1.0, 63.9995], [0.0, 64.0, 32.0], [2**-20, 0.25, 0.5]])
    las, headers = port_oracle.payload_las(xyz, np.arange(4, dtype=np.uint32), nodes, bounds)
    want = [[int(v / 0.001 + 0.5) for v in row] for row in xyz]
    assert las.tolist() == want
    assert headers["scale"][0] == 0.001 and headers["offset"][0].tolist() == [0.0, 0.0, 0.0]


# ---------------------------------------------------------------------------------------------------
# CUDA path through the C ABI
# ---------------------------------------------------------------------------------------------------
def gpu_case(name, shift, sampling="GRID_CENTER", tiling="ACCURATE", max_points=500, device_input=False):
    import schwarzwald_b200 as sw
    from schwarzwald_b200 import tiler as swt
    las, scale, offset, hmin, hmax = gio.las_input(name)
    cmin, cmax = gio.cubic(hmin, hmax)
    center = cmin + (cmax - cmin) / 2 if shift else None
    bmin, bmax = (cmin - center, cmax - center) if shift else (cmin, cmax)
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax, 12.0 if name == "small_object" else 250.0)
    t = swt.las_transform(scale, offset, hmin, hmax, center)
    with sw.GpuTiler(sampling, tiling, bmin, bmax, spacing, max_points_per_node=max_points, concurrency=2) as g:
        if device_input:
            import torch
            g.build_execution_graph_las(torch.from_numpy(las).cuda(), t)
        else:
            g.build_execution_graph_las(las, t)
        g.finalize()
        res = g.result()
        xyz = g.positions(len(las))
        pnts = g.payload_pnts()
        las_out, headers = g.payload_las()
    return dict(xyz=xyz, result=res, pnts=pnts, las_out=las_out, headers=headers)


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLDEN, ids=CASE_IDS)
def test_gpu_las_pipeline_equals_committed_reference_golden(g):
    """LAS records -> device positions -> tiling -> payloads, against digests of the reference's own code.
    swgpu_get_positions returns the positions after index_point's clamping; none of these clouds leaves the
    cubic bounds (the reader already clamped to the header), so they equal the reader's output."""
    check_against_golden(gpu_case(g["las"], g["shift"]), g)


@pytest.mark.gpu
@pytest.mark.parametrize("sampling,tiling", [("RANDOM_GRID", "FAST"), ("JITTERED", "ACCURATE"), ("MIN_DISTANCE", "FAST")])
@pytest.mark.parametrize("name", ["small_object", "continental", "aniso_negative"])
def test_gpu_las_pipeline_equals_oracle(port_oracle, name, sampling, tiling):
    from oracle import sworacle
    from schwarzwald_b200.tiler import SwgpuError
    try:
        want = gio.tiling_case(port_oracle, name, True, sampling=sampling, tiling=tiling, max_points=300)
    except sworacle.OracleFailure as e:  # small_object's coarse spacing: "Grids smaller than 16x16 ..." (JITTERED)
        with pytest.raises(SwgpuError) as err:
            gpu_case(name, True, sampling=sampling, tiling=tiling, max_points=300, device_input=True)
        assert err.value.code == e.code
        return
    got = gpu_case(name, True, sampling=sampling, tiling=tiling, max_points=300, device_input=True)
    assert np.array_equal(want["clamped"], got["xyz"])
    wt, wi = want["result"].canonical()
    gt, gi = got["result"].canonical()
    assert np.array_equal(wt[:, :3], gt[:, :3]) and np.array_equal(wi, gi)
    for key in ("pnts", "las_out"):
        a = gio.canonical_payload(want["result"], want[key])
        b = gio.canonical_payload(got["result"], got[key])
        assert a.dtype == b.dtype and np.array_equal(a.view(np.uint32), b.view(np.uint32)), key
    a = gio.canonical_headers(want["result"], want["headers"])
    b = gio.canonical_headers(got["result"], got["headers"])
    assert a.tobytes() == b.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 5, 4099])
def test_gpu_las_ragged_sizes(port_oracle, n):
    """the four-records-per-thread kernel and its tail"""
    import schwarzwald_b200 as sw
    from oracle import sworacle
    from schwarzwald_b200 import tiler as swt
    las, scale, offset, hmin, hmax = gio.las_input("utm_mm", n=max(n, 8))
    las = np.ascontiguousarray(las[4:4 + n] if n <= 4 else las[:n])
    cmin, cmax = gio.cubic(hmin, hmax)
    spacing = sw.spacing_from_diagonal_fraction(cmin, cmax)
    want_xyz = port_oracle.las_positions(las, sworacle.make_las_transform(scale, offset, hmin, hmax))
    want = port_oracle.tile(sworacle.make_params("GRID_CENTER", "ACCURATE", spacing, cmin, cmax), want_xyz)
    with sw.GpuTiler("GRID_CENTER", "ACCURATE", cmin, cmax, spacing) as g:
        g.build_execution_graph_las(las, swt.las_transform(scale, offset, hmin, hmax))
        g.finalize()
        got = g.result()
        assert np.array_equal(g.positions(n), want_xyz)
        assert np.array_equal(g.keys(n)[0], want.keys)
    wt, wi = want.canonical()
    gt, gi = got.canonical()
    assert np.array_equal(wt[:, :3], gt[:, :3]) and np.array_equal(wi, gi)


@pytest.mark.gpu
def test_gpu_payloads_after_double_input(port_oracle):
    """the payload calls also serve the classic PointBuffer (double) entry point, outliers included"""
    import schwarzwald_b200 as sw
    from oracle import sworacle
    rng = np.random.default_rng(11)
    xyz = np.round(rng.random((70_000, 3)) * [300.0, 280.0, 40.0] + [1000.0, -50.0, 5.0], 3)
    xyz[:9] += 500.0
    bmin, bmax = sw.cubic_bounds(xyz[9:].min(0), xyz[9:].max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    p = sworacle.make_params("RANDOM_GRID", "FAST", spacing, bmin, bmax, max_points_per_node=400, concurrency=2)
    want, clamped = port_oracle.tile(p, xyz, return_clamped=True)
    with sw.GpuTiler("RANDOM_GRID", "FAST", bmin, bmax, spacing, max_points_per_node=400, concurrency=2) as g:
        got = g.tile(xyz.copy())
        pnts = g.payload_pnts()
        las, headers = g.payload_las()
        assert np.array_equal(g.positions(len(xyz)), clamped)
    assert np.array_equal(want.canonical()[1], got.canonical()[1])
    w_pnts = port_oracle.payload_pnts(clamped, got.ids)
    w_las, w_headers = port_oracle.payload_las(clamped, got.ids, got.nodes, (bmin, bmax))
    assert np.array_equal(pnts.view(np.uint32), w_pnts.view(np.uint32))
    assert np.array_equal(las, w_las) and headers.tobytes() == w_headers.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("sampling", ["RANDOM_GRID", "GRID_CENTER"])
def test_gpu_payload_device_variants_equal_host_variants(sampling):
    """swgpu_get_payload_*_device write the same records into caller-owned device buffers.  RANDOM_GRID reads the
    positions through the sort permutation, the other strategies from the Morton-ordered copy."""
    import torch
    import schwarzwald_b200 as sw
    rng = np.random.default_rng(17)
    xyz = np.round(rng.random((90_000, 3)) * [120.0, 90.0, 30.0] + [10.0, 20.0, 5.0], 3)
    bmin, bmax = sw.cubic_bounds(xyz.min(0), xyz.max(0))
    spacing = sw.spacing_from_diagonal_fraction(bmin, bmax)
    with sw.GpuTiler(sampling, "FAST", bmin, bmax, spacing, max_points_per_node=600, concurrency=2) as g:
        g.tile(xyz.copy())
        _, ni = g.result_size()
        pnts, (las, headers) = g.payload_pnts(), g.payload_las()
        d_pnts = torch.empty((ni, 3), dtype=torch.float32, device="cuda")
        d_las = torch.empty((ni, 3), dtype=torch.int32, device="cuda")
        g.payload_pnts_device(d_pnts.data_ptr())
        d_headers = g.payload_las_device(d_las.data_ptr())
        torch.cuda.synchronize()
    assert np.array_equal(d_pnts.cpu().numpy().view(np.uint32), pnts.view(np.uint32))
    assert np.array_equal(d_las.cpu().numpy(), las) and d_headers.tobytes() == headers.tobytes()


def test_las_payload_round_trip_matches_the_reference_test(port_oracle):
    """Port of the expectation in the reference's test/TestLASPersistence.cpp:42-55,68-129: 4096 random points in
    the unit cube, stored per root octant through LASPersistence and read back (offset + X * scale), come back
    changed but by less than 0.001."""
    from oracle import sworacle
    rng = np.random.default_rng(0)
    xyz = rng.random((4096, 3))
    bounds = (np.zeros(3), np.ones(3))
    keys, _ = port_oracle.index_points(xyz, bounds)
    order = np.argsort(keys, kind="stable")
    octant = (keys[order] >> np.uint64(60)).astype(np.int64)
    nodes = np.zeros(8, sworacle.NODE_DTYPE)
    for o in range(8):
        sel = np.nonzero(octant == o)[0]
        # the reference test hands every octant's points to persist_points with the ROOT bounds
        # (get_bounds_from_morton_index(..., depth 0)): rows with zero levels
        nodes[o] = (0, 0, 0, sel[0] if len(sel) else 0, len(sel))
    las, headers = port_oracle.payload_las(xyz, order.astype(np.uint32), nodes, bounds)
    worst, changed = 0.0, False
    for o in range(8):
        first, count = int(nodes["first"][o]), int(nodes["count"][o])
        back = headers["offset"][o] + las[first:first + count] * headers["scale"][o]
        src = xyz[order[first:first + count]]
        changed |= bool((back != src).any())
        worst = max(worst, float(np.sqrt(((back - src) ** 2).sum(axis=1)).max()))
        assert headers["scale"][o] == 0.001  # diagonal sqrt(3) > 1
    assert changed and worst < 0.001


def test_las_positions_match_plain_ieee_arithmetic(port_oracle):
    """Independent of both C++ builds: offset + X * scale (two roundings), clamp, subtract the centre, round to
    float32 — spelled out in numpy float64 / float32."""
    from oracle import sworacle
    rng = np.random.default_rng(21)
    las = rng.integers(-2**31, 2**31 - 1, size=(30_000, 3), dtype=np.int64).astype(np.int32)
    las[::3] = rng.integers(0, 4_000_000, size=(10_000, 3))
    scale = np.array([0.001, 0.0125, 1e-4])
    offset = np.array([389000.5, 5705000.25, -3.0])
    hmin = np.array([389000.0, 5705000.0, -10.0])
    hmax = np.array([393000.75, 5745000.5, 380.0])
    center = np.array([391000.375, 5725000.25, 185.0])
    for shift in (False, True):
        t = sworacle.make_las_transform(scale, offset, hmin, hmax, center if shift else None)
        p = offset + las.astype(np.float64) * scale
        p = np.minimum(hmax, np.maximum(hmin, p))
        if shift:
            p = (p - center).astype(np.float32).astype(np.float64)
        assert np.array_equal(port_oracle.las_positions(las, t), p)
