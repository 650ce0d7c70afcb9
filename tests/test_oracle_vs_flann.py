"""CPU: the oracle's kd-tree (oracle/kdtree.cc, a restatement of FLANN's KDTreeSingleIndex as pcl::KdTreeFLANN configures it)
against a REAL FLANN build -- the copy OpenCV vendors (cv2.flann_Index, algorithm 4 = FLANN_INDEX_KDTREE_SINGLE, leaf_max_size 15,
L2 over the three coordinates in float32).  FLANN is the one third-party library of the reference's stack that exists in this
image, so this is the one place where the oracle is pinned to upstream code instead of to its own restatement
(DESIGN.md section 2, row 11 of the third-party table): exact k-NN, float32 squared distances accumulated as (dx^2 + dy^2) + dz^2,
ascending order.  FLANN leaves the order of exactly tied neighbours unspecified; rows with a tie inside the first k + 1 distances
are compared as sets.  Call sites pinned: kdtreeCornerLast / kdtreeSurfLast->nearestKSearch(pointSel, 1, ...)
(laserOdometry.cpp:302,390) and kdtreeCornerFromMap / kdtreeSurfFromMap->nearestKSearch(pointSel, 5, ...)
(laserMapping.cpp:584,652)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
if not hasattr(cv2, "flann_Index"):
    pytest.skip("this OpenCV build has no FLANN module", allow_module_level=True)

KDTREE_SINGLE = 4


def flann_knn(cloud_xyz, queries_xyz, k):
    index = cv2.flann_Index(np.ascontiguousarray(cloud_xyz, np.float32), dict(algorithm=KDTREE_SINGLE, leaf_max_size=15))
    idx, dist = index.knnSearch(np.ascontiguousarray(queries_xyz, np.float32), k, params=dict(checks=-1, eps=0.0, sorted=True))
    return idx.astype(np.int32), dist.astype(np.float32)


def compare(orc, cloud, queries, k):
    """returns (#rows compared index by index, #rows with ties compared as sets)"""
    oi, od = orc.KdTree(cloud).knn(queries, k)
    fi, fd = flann_knn(cloud[:, :3], queries[:, :3], k)
    assert od.dtype == np.float32 and np.array_equal(od, fd), "squared distances differ from FLANN's"
    # ties: equal consecutive distances inside the list, or the k-th distance shared with a point outside it
    kk = min(k + 1, cloud.shape[0])
    _, fd1 = flann_knn(cloud[:, :3], queries[:, :3], kk)
    tied = (np.diff(fd1, axis=1) == 0).any(axis=1)
    assert np.array_equal(oi[~tied], fi[~tied]), "neighbour lists differ from FLANN's"
    for r in np.nonzero(tied)[0]:
        d_o = np.sort(((cloud[oi[r], :3] - queries[r, :3]) ** 2).sum(1))
        d_f = np.sort(((cloud[fi[r], :3] - queries[r, :3]) ** 2).sum(1))
        assert np.allclose(d_o, d_f, rtol=0, atol=0)
    return int((~tied).sum()), int(tied.sum())


@pytest.mark.parametrize("n,k,scale", [(16, 1, 5.0), (17, 5, 5.0), (1000, 1, 20.0), (5000, 5, 20.0), (40000, 5, 60.0), (40000, 1, 0.5)])
def test_random_clouds(orc, n, k, scale):
    rng = np.random.default_rng(n + 31 * k)
    cloud = np.zeros((n, 4), np.float32); cloud[:, :3] = (rng.standard_normal((n, 3)) * scale).astype(np.float32)
    q = np.zeros((500, 4), np.float32); q[:, :3] = (rng.standard_normal((500, 3)) * scale).astype(np.float32)
    exact, tied = compare(orc, cloud, q, k)
    assert exact >= 450


def test_the_searches_of_the_registration_path(orc, synth):
    """the clouds and queries the two registration stages really search: scan k's sharp / flat points against scan k-1's
    less-sharp / less-flat clouds (k = 1), and a scan's feature clouds against a map accumulated from other scans (k = 5)"""
    ns, _, mr = synth.SENSORS["VLP-16"][:3]
    f = [orc.Features(synth.scan("VLP-16", k, n_az=900), ns, mr) for k in range(4)]
    rows = 0
    for k in range(1, 4):
        rows += compare(orc, f[k - 1].less_sharp, f[k].sharp, 1)[0]
        rows += compare(orc, f[k - 1].less_flat, f[k].flat, 1)[0]
    cmap = np.concatenate([f[k].less_sharp for k in (0, 1, 3)])
    smap = np.concatenate([f[k].less_flat for k in (0, 1, 3)])
    rows += compare(orc, cmap, f[2].less_sharp, 5)[0]
    rows += compare(orc, smap, f[2].less_flat, 5)[0]
    assert rows > 2000


def test_lattice_with_ties(orc):
    """many exactly equal distances: the distances still agree, the lists agree as sets"""
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    cloud = np.zeros((g.shape[0], 4), np.float32); cloud[:, :3] = g
    q = np.zeros((50, 4), np.float32); q[:, :3] = np.random.default_rng(5).integers(0, 12, (50, 3)).astype(np.float32) + 0.5
    exact, tied = compare(orc, cloud, q, 5)
    assert tied > 0
