"""Synthetic structured meshes and function-space tables (the INPUTS of the assembly path) for tests and benchmarks.

The reference builds these with Mesh.Parallelepiped / GetHighOrderMesh (Florence/MeshGeneration/Mesh.py:5843, :1339) and
FunctionSpace (Florence/FunctionSpace/FunctionSpace.py:81-168).  Its high-order mesher is a per-element Python loop that
cannot produce the 8M-64M element meshes of the benchmark configs, so this module generates straight-sided
structured meshes directly with vectorised torch ops (on the GPU when `device` is a CUDA device).  Node ordering is our own
(lexicographic tensor-product / Kuhn simplices) -- the assembly path only needs elements, points and tables to be consistent;
tests/test_mesh_tables.py checks the tables against the reference's FunctionSpace output through ordering-invariant quantities.

Tables have the reference's shapes: Bases (npe x ng), Jm (ndim x npe x ng), AllGauss (ng x 1).
"""
import itertools

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ 1-D building blocks
def gauss_legendre(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


def gll_nodes(p):
    """Gauss-Lobatto-Legendre points on [-1, 1] (p+1 of them); equally spaced for p <= 2."""
    if p == 1:
        return np.array([-1.0, 1.0])
    c = np.zeros(p + 1)
    c[p] = 1.0
    dP = np.polynomial.legendre.legder(c)
    return np.concatenate([[-1.0], np.sort(np.polynomial.legendre.legroots(dP)), [1.0]])


def lagrange_1d(nodes, x):
    """Values L (n x m) and derivatives dL (n x m) of the Lagrange basis on `nodes` at points x."""
    n, m = len(nodes), len(x)
    L = np.ones((n, m))
    dL = np.zeros((n, m))
    for i in range(n):
        for j in range(n):
            if j != i:
                L[i] *= (x - nodes[j]) / (nodes[i] - nodes[j])
        for k in range(n):
            if k == i:
                continue
            term = np.ones(m) / (nodes[i] - nodes[k])
            for j in range(n):
                if j != i and j != k:
                    term *= (x - nodes[j]) / (nodes[i] - nodes[j])
            dL[i] += term
    return L, dL


# ------------------------------------------------------------------------------------------------ tables
def hex_tables(p, ngauss_1d=None):
    """Tensor-product Lagrange (GLL nodes) hexahedron, Gauss-Legendre (p+1)^3 rule (norder = C+2, VariationalPrinciple.py:53-72).
    Local node a = i + (p+1) (j + (p+1) k), x fastest; Gauss point g likewise."""
    n1 = p + 1
    q = ngauss_1d or p + 1
    xg, wg = gauss_legendre(q)
    L, dL = lagrange_1d(gll_nodes(p), xg)
    npe, ng = n1 ** 3, q ** 3
    Bases = np.einsum("kc,jb,ia->kjicba", L, L, L).reshape(npe, ng)
    Jm = np.zeros((3, npe, ng))
    Jm[0] = np.einsum("kc,jb,ia->kjicba", L, L, dL).reshape(npe, ng)
    Jm[1] = np.einsum("kc,jb,ia->kjicba", L, dL, L).reshape(npe, ng)
    Jm[2] = np.einsum("kc,jb,ia->kjicba", dL, L, L).reshape(npe, ng)
    AllGauss = np.einsum("c,b,a->cba", wg, wg, wg).reshape(ng, 1)
    return np.ascontiguousarray(Bases), np.ascontiguousarray(Jm), np.ascontiguousarray(AllGauss)


def quad_tables(p, ngauss_1d=None):
    n1 = p + 1
    q = ngauss_1d or p + 1
    xg, wg = gauss_legendre(q)
    L, dL = lagrange_1d(gll_nodes(p), xg)
    npe, ng = n1 ** 2, q ** 2
    Bases = np.einsum("jb,ia->jiba", L, L).reshape(npe, ng)
    Jm = np.zeros((2, npe, ng))
    Jm[0] = np.einsum("jb,ia->jiba", L, dL).reshape(npe, ng)
    Jm[1] = np.einsum("jb,ia->jiba", dL, L).reshape(npe, ng)
    AllGauss = np.einsum("b,a->ba", wg, wg).reshape(ng, 1)
    return np.ascontiguousarray(Bases), np.ascontiguousarray(Jm), np.ascontiguousarray(AllGauss)


def _gauss_jacobi_01(n, alpha):
    """n-point Gauss rule on [0,1] for the weight (1-x)^alpha (Golub-Welsch on the Jacobi recurrence)."""
    from scipy.special import roots_jacobi
    x, w = roots_jacobi(n, alpha, 0.0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)


def simplex_rule(ndim, n):
    """Collapsed-coordinate (Stroud conical product) rule with n^ndim points on the unit simplex, exact to degree 2n-1."""
    if ndim == 2:
        x1, w1 = _gauss_jacobi_01(n, 1)
        x2, w2 = _gauss_jacobi_01(n, 0)
        pts = [(a, b * (1 - a)) for a in x1 for b in x2]
        wts = [wa * wb for wa in w1 for wb in w2]
    else:
        x1, w1 = _gauss_jacobi_01(n, 2)
        x2, w2 = _gauss_jacobi_01(n, 1)
        x3, w3 = _gauss_jacobi_01(n, 0)
        pts = [(a, b * (1 - a), c * (1 - a) * (1 - b)) for a in x1 for b in x2 for c in x3]
        wts = [wa * wb * wc for wa in w1 for wb in w2 for wc in w3]
    return np.array(pts), np.array(wts)


TET10_EDGES = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))
TRI6_EDGES = ((0, 1), (1, 2), (0, 2))


def simplex_tables(ndim, p, npoints_1d=None):
    """P1/P2 Lagrange simplex (vertices, then edge midpoints in TET10_EDGES / TRI6_EDGES order) on the unit simplex.
    Default rule: 2^ndim conical-product points for p=2 (8 points for tet10, the count the reference's degree-2(p-1) table
    gives -- SURVEY.md 8), 1 point for p=1."""
    n = npoints_1d or (1 if p == 1 else 2)
    pts, wts = simplex_rule(ndim, n)
    ng = len(wts)
    lam = np.concatenate([(1.0 - pts.sum(1))[:, None], pts], axis=1)  # (ng, ndim+1)
    dlam = np.concatenate([-np.ones((1, ndim)), np.eye(ndim)], axis=0)  # (ndim+1, ndim)
    nv = ndim + 1
    edges = TET10_EDGES if ndim == 3 else TRI6_EDGES
    if p == 1:
        N = lam.T.copy()
        dN = np.repeat(dlam[:, :, None], ng, axis=2)
    elif p == 2:
        npe = nv + len(edges)
        N = np.zeros((npe, ng))
        dN = np.zeros((npe, ndim, ng))
        for i in range(nv):
            N[i] = lam[:, i] * (2 * lam[:, i] - 1)
            dN[i] = dlam[i][:, None] * (4 * lam[:, i] - 1)[None, :]
        for k, (i, j) in enumerate(edges):
            N[nv + k] = 4 * lam[:, i] * lam[:, j]
            dN[nv + k] = 4 * (dlam[i][:, None] * lam[:, j][None, :] + dlam[j][:, None] * lam[:, i][None, :])
    else:
        raise NotImplementedError("simplex tables are provided for p = 1, 2")
    Jm = np.ascontiguousarray(np.transpose(dN, (1, 0, 2)))
    return np.ascontiguousarray(N), Jm, wts.reshape(ng, 1).copy()


def tables(element_type, p, **kw):
    if element_type == "hex":
        return hex_tables(p, **kw)
    if element_type == "quad":
        return quad_tables(p, **kw)
    if element_type == "tet":
        return simplex_tables(3, p, **kw)
    if element_type == "tri":
        return simplex_tables(2, p, **kw)
    raise ValueError(element_type)


# ------------------------------------------------------------------------------------------------ meshes
def _axis_coords(n, p, length, device):
    ref = torch.as_tensor((gll_nodes(p) + 1.0) / 2.0, dtype=torch.float64, device=device)  # (p+1) in [0,1]
    e = torch.arange(n, dtype=torch.float64, device=device)
    x = (e[:, None] + ref[None, :p]).reshape(-1)
    x = torch.cat([x, torch.tensor([float(n)], dtype=torch.float64, device=device)])
    return x * (length / n)


def box_hex_mesh(nx, ny, nz, p=1, lengths=(1.0, 1.0, 1.0), device="cpu", z_offset_elems=0, nz_total=None):
    """Straight-sided order-p hexahedra on a box.  Global node (I,J,K) -> I + NX (J + NY K) with NX = p nx + 1; elements
    ordered x fastest.  Returns points (nnode x 3) float64 and elements (nelem x (p+1)^3) int64.
    z_offset_elems / nz_total describe a z-slab of a taller box (used by the multi-GPU partitioner)."""
    dev = torch.device(device)
    nzt = nz_total or nz
    NX, NY, NZ = p * nx + 1, p * ny + 1, p * nz + 1
    xs = _axis_coords(nx, p, lengths[0], dev)
    ys = _axis_coords(ny, p, lengths[1], dev)
    zs_full = _axis_coords(nzt, p, lengths[2], dev)
    zs = zs_full[p * z_offset_elems: p * z_offset_elems + NZ]
    Z, Y, X = torch.meshgrid(zs, ys, xs, indexing="ij")
    points = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], dim=1).contiguous()
    ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
    base = (p * ar(nx))[None, None, :] + NX * ((p * ar(ny))[None, :, None] + NY * (p * ar(nz))[:, None, None])
    loc = ar(p + 1)[None, None, :] + NX * (ar(p + 1)[None, :, None] + NY * ar(p + 1)[:, None, None])
    elements = (base.reshape(-1, 1) + loc.reshape(1, -1)).contiguous()
    return points, elements


def rect_quad_mesh(nx, ny, p=1, lengths=(1.0, 1.0), device="cpu"):
    dev = torch.device(device)
    NX = p * nx + 1
    xs = _axis_coords(nx, p, lengths[0], dev)
    ys = _axis_coords(ny, p, lengths[1], dev)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    points = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=1).contiguous()
    ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
    base = (p * ar(nx))[None, :] + NX * (p * ar(ny))[:, None]
    loc = ar(p + 1)[None, :] + NX * ar(p + 1)[:, None]
    return points, (base.reshape(-1, 1) + loc.reshape(1, -1)).contiguous()


def box_tet_mesh(nx, ny, nz, p=1, lengths=(1.0, 1.0, 1.0), device="cpu"):
    """Kuhn split of each hexahedron into 6 positively oriented tetrahedra (6 per hex as the reference's
    ConvertHexesToTets, Mesh.py:8757-8782); p=2 adds the edge midpoints, which are the points of the (2n+1)^3 fine grid."""
    if p not in (1, 2):
        raise NotImplementedError
    dev = torch.device(device)
    NX, NY, NZ = p * nx + 1, p * ny + 1, p * nz + 1
    xs = torch.linspace(0.0, lengths[0], NX, dtype=torch.float64, device=dev)
    ys = torch.linspace(0.0, lengths[1], NY, dtype=torch.float64, device=dev)
    zs = torch.linspace(0.0, lengths[2], NZ, dtype=torch.float64, device=dev)
    Z, Y, X = torch.meshgrid(zs, ys, xs, indexing="ij")
    points = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], dim=1).contiguous()
    # local vertex offsets (in hex-vertex units) of the 6 Kuhn tets
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            w = v[-1].copy()
            w[ax] += 1
            v.append(w)
        v = np.array(v)
        if np.linalg.det((v[1:] - v[0]).astype(float)) < 0:
            v[[1, 2]] = v[[2, 1]]
        tets.append(v)
    tets = np.array(tets)  # (6, 4, 3)
    if p == 2:
        mids = np.array([[tets[t, i] + tets[t, j] for (i, j) in TET10_EDGES] for t in range(6)])  # fine-grid offsets (sum = 2*mid)
        loc = np.concatenate([2 * tets, mids], axis=1)  # (6, 10, 3) in fine-grid units
    else:
        loc = tets
    loc = torch.as_tensor(loc, dtype=torch.int64, device=dev)
    loc_flat = loc[..., 0] + NX * (loc[..., 1] + NY * loc[..., 2])  # (6, npe)
    ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
    base = (p * ar(nx))[None, None, :] + NX * ((p * ar(ny))[None, :, None] + NY * (p * ar(nz))[:, None, None])
    elements = (base.reshape(-1, 1, 1) + loc_flat[None, :, :]).reshape(-1, loc_flat.shape[1]).contiguous()
    return points, elements


def rect_tri_mesh(nx, ny, p=1, lengths=(1.0, 1.0), device="cpu"):
    if p not in (1, 2):
        raise NotImplementedError
    dev = torch.device(device)
    NX, NY = p * nx + 1, p * ny + 1
    xs = torch.linspace(0.0, lengths[0], NX, dtype=torch.float64, device=dev)
    ys = torch.linspace(0.0, lengths[1], NY, dtype=torch.float64, device=dev)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    points = torch.stack([X.reshape(-1), Y.reshape(-1)], dim=1).contiguous()
    tris = np.array([[[0, 0], [1, 0], [1, 1]], [[0, 0], [1, 1], [0, 1]]])
    if p == 2:
        mids = np.array([[tris[t, i] + tris[t, j] for (i, j) in TRI6_EDGES] for t in range(2)])
        loc = np.concatenate([2 * tris, mids], axis=1)
    else:
        loc = tris
    loc = torch.as_tensor(loc, dtype=torch.int64, device=dev)
    loc_flat = loc[..., 0] + NX * loc[..., 1]
    ar = lambda n: torch.arange(n, dtype=torch.int64, device=dev)
    base = (p * ar(nx))[None, :] + NX * (p * ar(ny))[:, None]
    elements = (base.reshape(-1, 1, 1) + loc_flat[None, :, :]).reshape(-1, loc_flat.shape[1]).contiguous()
    return points, elements


def make_mesh(element_type, n, p, lengths=None, device="cpu"):
    n = (n,) * (3 if element_type in ("hex", "tet") else 2) if isinstance(n, int) else tuple(n)
    if element_type == "hex":
        return box_hex_mesh(*n, p=p, lengths=lengths or (1.0, 1.0, 1.0), device=device)
    if element_type == "tet":
        return box_tet_mesh(*n, p=p, lengths=lengths or (1.0, 1.0, 1.0), device=device)
    if element_type == "quad":
        return rect_quad_mesh(*n, p=p, lengths=lengths or (1.0, 1.0), device=device)
    if element_type == "tri":
        return rect_tri_mesh(*n, p=p, lengths=lengths or (1.0, 1.0), device=device)
    raise ValueError(element_type)


def boundary_nodes(points, tol=1e-12):
    """Sorted ids of the nodes on the surface of an axis-aligned box mesh -- np.unique(mesh.faces) (np.unique(mesh.edges) in 2-D) of
    the reference's Mesh for these structured boxes (ExplicitPenaltyContactFormulation.py:157-161 consumes exactly that list)."""
    is_t = isinstance(points, torch.Tensor)
    P = points if is_t else torch.as_tensor(np.asarray(points))
    lo, hi = P.min(0).values, P.max(0).values
    on = ((P - lo).abs() <= tol * (1 + hi.abs())) | ((P - hi).abs() <= tol * (1 + hi.abs()))
    ids = torch.nonzero(on.any(1)).reshape(-1)
    return ids if is_t else ids.numpy()


def perturbed_state(points, h, amplitude=0.02, seed=0):
    """Eulerx = X + amplitude*h*U(-1,1) per node, generated on the tensor's device with a fixed seed."""
    gen = torch.Generator(device=points.device)
    gen.manual_seed(seed)
    r = torch.rand(points.shape, dtype=torch.float64, device=points.device, generator=gen)
    return points + amplitude * h * (2.0 * r - 1.0)
