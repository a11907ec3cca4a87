"""Shared set-up for the parity tests: the same initial particles and fields in the CPU oracle
and in the CUDA path (through the C ABI)."""
import numpy as np

from oracle import oracle as O


def make_pair(n=(10, 10, 10), Np=100, T=0.1, scheme=0, curl_sign=+1, rtol=1e-12, precond=0, seed_fields=None, d=(0.5, 0.5, 0.5), dt=1.5,
              species=((-1.0, 1.0, 1.0),)):
    """Returns (oracle, gpu simulation) holding identical state.  Particles come from the
    reference's default-seeded mt19937 stream (oracle.set_particles_maxwell)."""
    import xpic_b200 as X

    o = O.Oracle(n, d=d, dt=dt, curl_sign=curl_sign)
    s = X.Simulation(n, d=d, dt=dt, scheme=scheme, curl_sign=curl_sign, track_ids=True)
    for (q, m, dens) in species:
        sid = o.add_species(q=q, m=m, n=dens, Np=Np)
        o.set_particles_maxwell(sid, T=T, tov=True)
        gs = s.add_species(q=q, m=m, n=dens, Np=Np)
        assert gs == sid
        pts, ids = o.get_particles(sid)
        assert s.add_particles(sid, pts, ids) == len(ids)
    for which in (0, 1):
        o.solver_set(which, rtol, 1e-50, 1000, 30)
        s.solver_set(which, rtol, 1e-50, 1000, 30, precond)
    if seed_fields is not None:
        rng = np.random.default_rng(seed_fields)
        for name, amp in (("E", 0.02), ("B", 0.05)):
            f = amp * rng.standard_normal(o.n3)
            o.set_field(name, f)
            s.set_field(name, f)
    return o, s


def by_id(pts, ids):
    order = np.argsort(ids, kind="stable")
    return pts[order], ids[order]


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def csr_to_stencil(o, table):
    """Oracle CSR of L -> coef[k][node] in the product's fixed-offset layout."""
    import scipy.sparse as sp

    rp, col, val = o.csr(0)
    A = sp.csr_matrix((val, col, rp), shape=(o.n3, o.n3))
    nx, ny, nz = o.n
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    node = ((z * ny + y) * nx + x).reshape(-1)
    coef = np.zeros((len(table), node.size))
    for k, (c1, c2, dx, dy, dz) in enumerate(table):
        rows = node * 3 + c1
        cn = ((((z + dz) % nz) * ny + ((y + dy) % ny)) * nx + ((x + dx) % nx)).reshape(-1)
        cols = cn * 3 + c2
        coef[k] = np.asarray(A[rows, cols]).reshape(-1)
    return coef


def csr_to_stencil_open(o, table):
    """As csr_to_stencil for a box whose z boundary is open.  Returns (coef, exists): exists[k, node] is False where the
    column of slot k lies outside the box -- the reference has no such matrix entry; the product keeps whatever the
    particles deposited there, it only ever multiplies the zero ghost values of x."""
    import scipy.sparse as sp

    rp, col, val = o.csr(0)
    A = sp.csr_matrix((val, col, rp), shape=(o.n3, o.n3))
    nx, ny, nz = o.n
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    node = ((z * ny + y) * nx + x).reshape(-1)
    coef = np.zeros((len(table), node.size))
    exists = np.zeros((len(table), node.size), dtype=bool)
    for k, (c1, c2, dx, dy, dz) in enumerate(table):
        rows = node * 3 + c1
        zc = (z + dz).reshape(-1)
        inside = (zc >= 0) & (zc < nz)
        cn = ((np.clip(zc, 0, nz - 1).reshape(z.shape) * ny + ((y + dy) % ny)) * nx + ((x + dx) % nx)).reshape(-1)
        cols = cn * 3 + c2
        v = np.asarray(A[rows, cols]).reshape(-1)
        coef[k] = np.where(inside, v, 0.0)
        exists[k] = inside
    return coef, exists


def spline2(s):
    s = np.abs(s)
    return np.where(s <= 0.5, 0.75 - s * s, np.where(s < 1.5, 0.5 * (1.5 - s) ** 2, 0.0))


def momentum_qe(pts, E, n, d, q_np):
    """QE of MomentumConservation::calculate (momentum_conservation.cpp:84-117): sum over particles of
    q / Np * E at the particle with the global 2nd-order Shape (electric(): No No Sh per component)."""
    nx, ny, nz = n
    Eg = E.reshape(nz, ny, nx, 3)
    p = pts[:, :3] / np.array(d)
    start = np.round(p - 1.5).astype(int)  # np.round is half-to-even; exact ties do not occur for these particles
    out = np.zeros(3)
    for k in range(4):
        for j in range(4):
            for i in range(4):
                g = start + np.array([i, j, k])
                no = [spline2(p[:, a] - g[:, a]) for a in range(3)]
                sh = [spline2(p[:, a] - (g[:, a] + 0.5)) for a in range(3)]
                idx = (g[:, 2] % nz, g[:, 1] % ny, g[:, 0] % nx)
                out[0] += np.sum(Eg[idx + (0,)] * (no[2] * no[1] * sh[0]))
                out[1] += np.sum(Eg[idx + (1,)] * (no[2] * sh[1] * no[0]))
                out[2] += np.sum(Eg[idx + (2,)] * (sh[2] * no[1] * no[0]))
    return q_np * out


def cell_density(pts, n, d, n_np):
    """DistributionMoment::collect with the "density" moment (src/diagnostics/distribution_moment.cpp:131-210):
    cell-centred, 1st-order form factor, two nodes per axis from round(p - 1), weight n / Np."""
    nx, ny, nz = n
    rho = np.zeros((nz, ny, nx))
    p = pts[:, :3] / np.array(d)
    start = np.round(p - 1.0).astype(int)

    def s1(s):
        s = np.abs(s)
        return np.where(s <= 1.0, 1.0 - s, 0.0)

    for k in range(2):
        wz = s1(p[:, 2] - (start[:, 2] + k + 0.5))
        for j in range(2):
            wy = s1(p[:, 1] - (start[:, 1] + j + 0.5))
            for i in range(2):
                wx = s1(p[:, 0] - (start[:, 0] + i + 0.5))
                np.add.at(rho, ((start[:, 2] + k) % nz, (start[:, 1] + j) % ny, (start[:, 0] + i) % nx), n_np * wx * wy * wz)
    return rho


def cell_moment(pts, n, d, n_np, name, q=-1.0, m=1.0, start=(0, 0, 0), size=None):
    """DistributionMoment::collect for any moment of src/diagnostics/distribution_moment.cpp:212-313 and a region:
    (nz, ny, nx, components) array; particles whose cell lies outside the region do not contribute, deposits outside the
    region are dropped unless the region spans the whole (periodic) axis."""
    nx, ny, nz = n
    size = tuple(n) if size is None else tuple(size)
    v = pts[:, 3:]
    if name == "density":
        mv = np.ones((len(pts), 1))
    elif name == "current":
        mv = q * v
    else:
        w = v.copy()
        if name.endswith("cyl"):
            x, y = pts[:, 0] - 0.5 * nx * d[0], pts[:, 1] - 0.5 * ny * d[1]
            r = np.hypot(x, y)
            ok = r > 0
            w[ok, 0] = (x[ok] * v[ok, 0] + y[ok] * v[ok, 1]) / r[ok]
            w[ok, 1] = (-y[ok] * v[ok, 0] + x[ok] * v[ok, 1]) / r[ok]
        if "diag" in name:
            mv = m * w * w
        else:
            mv = m * np.stack([w[:, 0] * w[:, 0], w[:, 0] * w[:, 1], w[:, 0] * w[:, 2], w[:, 1] * w[:, 1], w[:, 1] * w[:, 2], w[:, 2] * w[:, 2]], axis=1)
    out = np.zeros((nz, ny, nx, mv.shape[1]))
    p = pts[:, :3] / np.array(d)
    cell = np.floor(p).astype(int)
    inside = np.ones(len(pts), dtype=bool)
    for a in range(3):
        inside &= (cell[:, a] >= start[a]) & (cell[:, a] < start[a] + size[a])
    st = np.round(p - 1.0).astype(int)

    def s1(s):
        s = np.abs(s)
        return np.where(s <= 1.0, 1.0 - s, 0.0)

    for k in range(2):
        wz = s1(p[:, 2] - (st[:, 2] + k + 0.5))
        for j in range(2):
            wy = s1(p[:, 1] - (st[:, 1] + j + 0.5))
            for i in range(2):
                wx = s1(p[:, 0] - (st[:, 0] + i + 0.5))
                c = st + np.array([i, j, k])
                keep = inside.copy()
                for a in range(3):
                    if size[a] != n[a]:
                        keep &= (c[:, a] >= start[a]) & (c[:, a] < start[a] + size[a])
                idx = (c[keep, 2] % nz, c[keep, 1] % ny, c[keep, 0] % nx)
                np.add.at(out, idx, (n_np * wx * wy * wz)[keep, None] * mv[keep])
    return out
